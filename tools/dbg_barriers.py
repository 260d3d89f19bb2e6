import sys, time; sys.path.insert(0,'/root/repo')
import torch
from src.shoeprint_image_retrieval import engine, synth
gallery = synth.make_gallery(11, 5, 4, 16, 12)
probes, pairs = synth.make_probes(12, gallery, 4, min_frac=1.0)
torch.cuda.synchronize()
t0=time.time()
try:
    ranks, scores, _ = engine.compare(probes, gallery, pairs, None, None, precision=sys.argv[1] if len(sys.argv)>1 else "fp16x3")
    torch.cuda.synchronize()
    print(ranks)
except Exception as e:
    print("ERR after", time.time()-t0, str(e)[:100])
