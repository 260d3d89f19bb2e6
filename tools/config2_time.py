import sys, types
sys.path.insert(0, "/root/repo")
import __graft_entry__ as ge
ge.build()
import bench
args = types.SimpleNamespace(precision="fp16_refine")
print(bench.config0_block(args, 2))
