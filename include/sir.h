/*
 * sir.h -- C ABI of libsir.so, the B200 (sm_100a) matching hot path of
 * shoeprint-image-retrieval:  feature maps -> probe x gallery normalised cross-correlation
 * (max over offsets, rotations, scales) -> rank / top-k.
 *
 * The reference has no FFI layer: its boundary is the Python API of
 * src/shoeprint_image_retrieval/{similarity,network,parse_results}.py.  The
 * Python modules of this repository keep that API and reach the GPU only through the entry
 * points below (ctypes; see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, a negative SIR_E_* code otherwise; the message is
 *     available from sir_last_error() (thread local);
 *   - pointers named d_* are DEVICE pointers owned by the caller (PyTorch tensors in this
 *     repository); h_* are host pointers; nothing is allocated or freed behind the caller's back:
 *     calls that need scratch take a caller workspace sized by a *_workspace_bytes query;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, not synchronised;
 *   - feature maps are float32, row major [count][C][h][w]; a call handles maps of ONE shape,
 *     the host groups ragged inputs by shape (dataloader.py never pads);
 *   - "cropped" means the reference's [:, 2:-2, 2:-2] (similarity.py:92-93): Hp = hg-4 etc.
 */
#ifndef SIR_H_
#define SIR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIR_OK 0
#define SIR_E_ARG (-1)     /* invalid argument / unsupported shape */
#define SIR_E_CUDA (-2)    /* CUDA runtime or driver error */
#define SIR_E_DEVICE (-3)  /* not an sm_100 device */

/* precision modes of sir_ncc_scores */
#define SIR_PREC_FP16X3 0  /* tcgen05, fp16 hi/lo split operands, 3 MMAs per K step (parity grade, default) */
#define SIR_PREC_FP16X1 1  /* tcgen05, fp16 operands, 1 MMA per K step (fast; ~1e-4 relative worst case) */
#define SIR_PREC_FP32_SIMT 2 /* CUDA cores, fp32 FMA (exact-order independent check path) */
#define SIR_PREC_FP16_FP8C 3 /* tcgen05, hi*hi in fp16 + the two correction products in fp8 e4m3 (2/3 of the
                                FP16X3 tensor cycles, ~2^-14 relative per product); own entry points *_fp8c */
#define SIR_PREC_FP16_REFINE 4 /* tcgen05 screening with plain fp16 operands (1 MMA per K step) + exact float32
                                  re-evaluation of the positions that can hold the maximum (parity grade, default);
                                  entry points sir_ncc_screen + sir_ncc_refine */

const char* sir_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int sir_abi_version(void);
/* SM count / compute capability of the current device; SIR_E_DEVICE when it is not sm_100. */
int sir_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------ gallery side (K5 + K6)
 * Replaces, for the gallery operand, similarity.py:92 (crop), :49 (zero mean) and prepares the
 * operands of :53-55 (numerator) and :57-65 (window energy).
 *
 * d_gallery [G][C][hg][wg] f32  ->  d_ghi, d_glo [G][C][Hp][WP] f16, WP = sir_gallery_pitch(Wp)
 * (rows padded with zeros to a multiple of 8 cells so TMA can address them): (g - mean) * 2^e split
 * as hi = f16(x), lo = f16(x - hi);  d_gexp [G][C] int32: the exponent e (chosen so the channel's
 * max |value| lands in [2^9, 2^10), exact power of two so nothing is rounded by the scaling);
 * d_gz [G][C][Hp][Wp] f32 or NULL: the zero-meaned, UNscaled map (operand of the SIMT path). */
int sir_gallery_pitch(int Wp);
int sir_gallery_pack(const float* d_gallery, int G, int C, int hg, int wg,
                     uint16_t* d_ghi, uint16_t* d_glo, int32_t* d_gexp, float* d_gz, void* stream);

/* Window inverse norm for one template shape (Hm x Wm, cropped size), similarity.py:57-65,68:
 * d_rnorm [G][C][Hp*Wp] f32 = 1/sqrt(D) of the SCALED map (hi+lo), 0 where D <= 0, with
 * D = S2 - S1^2/(Hm*Wm) over the zero padded window anchored at (Hm/2, Wm/2), sums in float64.
 * If d_gz != NULL the unscaled map is used instead (for the SIMT path) and d_ghi/d_glo ignored. */
int sir_gallery_window_rnorm(const uint16_t* d_ghi, const uint16_t* d_glo, const float* d_gz,
                             int G, int C, int Hp, int Wp, int Hm, int Wm, float* d_rnorm, void* stream);

/* The same for `nshapes` template shapes in one pass over the gallery (the summed-area tables are built
 * once per channel): h_hm/h_wm host arrays of shapes, h_out host array of DEVICE pointers to the tables. */
int sir_gallery_window_rnorm_multi(const uint16_t* d_ghi, const uint16_t* d_glo, int G, int C, int Hp, int Wp, int nshapes,
                                   const int* h_hm, const int* h_wm, float* const* h_out, void* stream);

/* ------------------------------------------------------------------ probe side (K4 + K5)
 * Variant generation, similarity.py:262-276 (Pillow rotate nearest / resize bicubic on mode "F").
 * d_in [N][C][h][w] f32 -> d_out [N][C][h2][w2] f32.
 * rotate: angle in degrees (counter clockwise), same size, zero fill, 16.16 fixed point walk.
 * resize: bicubic a=-0.5, horizontal pass first, double accumulation; d_tmp [N][C][h][w2] f32
 * scratch (may be NULL when w2 == w or h2 == h); d_ws: caller-owned device workspace (16-byte aligned) of at least
 * sir_variant_resize_workspace_bytes(h, w, h2, w2) bytes for the tap tables of the passes (nothing is allocated
 * inside; the tables are written stream ordered, so the workspace may be reused by the next call on the same stream). */
int sir_variant_rotate(const float* d_in, int N, int C, int h, int w, double angle, float* d_out, void* stream);
size_t sir_variant_resize_workspace_bytes(int h, int w, int h2, int w2);
int sir_variant_resize(const float* d_in, int N, int C, int h, int w, int h2, int w2,
                       float* d_out, float* d_tmp, void* d_ws, size_t ws_bytes, void* stream);

/* Loader resize on the device (dataloader.py:231-237, Image.resize(size, LANCZOS) of the cropped 8-bit prints): Pillow's
 * 8 bits-per-channel resampler, bit exact (LANCZOS support 3, coefficients rounded to 22 fractional bits, horizontal
 * pass first into an 8-bit intermediate).  d_in uint8 [N][h][w][ch] (ch = 1 grayscale, 3 RGB) -> d_out [N][h2][w2][ch];
 * d_tmp [N][h][w2][ch] scratch (may be NULL when only one axis changes); d_ws: caller workspace of
 * sir_image_resize_workspace_bytes(h, w, h2, w2) bytes, 16-byte aligned. */
size_t sir_image_resize_workspace_bytes(int h, int w, int h2, int w2);
int sir_image_resize_lanczos(const uint8_t* d_in, int N, int h, int w, int ch, int h2, int w2, uint8_t* d_out, uint8_t* d_tmp,
                             void* d_ws, size_t ws_bytes, void* stream);

/* [N][C][h][w] -> [N][C][w][h].  Scores are invariant under transposing probe and gallery maps alike; the
 * host uses this to present the correlation kernel with the orientation that pads less. */
int sir_maps_transpose(const float* d_in, int N, int C, int h, int w, float* d_out, void* stream);

/* K-padded length of a packed template of cropped shape Hm x Wm: taps are ordered row by row,
 * each row padded to a multiple of 8 taps, the total to a multiple of 32 (one TMA box). */
int sir_template_kpad(int Hm, int Wm);

/* Template operand, similarity.py:92 (crop), :48 (zero mean), :67 (energy E):
 * d_maps [N][C][h][w] f32 (Hm = h-4, Wm = w-4) -> column col0+n of
 * d_thi, d_tlo [C][ncols_alloc][Kpad] f16 = (t - mean)/sqrt(E) * 2^10 split hi/lo in the padded
 * tap order above (all zero when E == 0: the reference maps the resulting non-finite NCC to 0,
 * similarity.py:70), and d_t32 [C][ncols_alloc][Hm*Wm] f32 or NULL = (t - mean)/sqrt(E). */
int sir_template_pack(const float* d_maps, int N, int C, int h, int w, int col0, int ncols_alloc,
                      uint16_t* d_thi, uint16_t* d_tlo, float* d_t32, void* stream);

/* ------------------------------------------------------------------ correlation (K7)
 * Replaces the Q*G*V calls of get_similarity (similarity.py:75-108, 357-367) for one template
 * shape: for every gallery g and packed column n,
 *     s = (1/C) max_{y,x} sum_c rnorm[g][c][y,x] * sum_{u,v} t_n,c[u,v] g_c[y+u-Hm/2, x+v-Wm/2]
 * and d_scores[col2probe[n] * score_ld + g0 + g] = max(old, s)  (float32, callers zero it first:
 * the reference floors at 0, similarity.py:355).  The correlation surface never leaves the SM.
 * Workspace: none.  precision: SIR_PREC_*. */
int sir_ncc_scores(const uint16_t* d_ghi, const uint16_t* d_glo, const int32_t* d_gexp, const float* d_gz,
                   const float* d_rnorm, int G, int C, int Hp, int Wp,
                   const uint16_t* d_thi, const uint16_t* d_tlo, const float* d_t32,
                   int ncols, int ncols_alloc, int Hm, int Wm,
                   const int32_t* d_col2probe, float* d_scores, int score_ld, int g0,
                   int precision, void* stream);

/* normxcorr (similarity.py:26-72) for one channel pair, surface written out (debug/helper path only):
 * d_gz [Hp][Wp] zero-meaned image, d_rnorm its window inverse norm for Hm x Wm, d_t32 [Hm][Wm] the packed
 * template (t - mean)/sqrt(E);  d_out [Hp][Wp] f32. */
int sir_ncc_surface(const float* d_gz, const float* d_rnorm, int Hp, int Wp, const float* d_t32, int Hm, int Wm,
                    float* d_out, void* stream);

/* fp8-corrected variant (SIR_PREC_FP16_FP8C).  Same quantity as sir_ncc_scores; operands:
 *   gallery : d_ghi (as above) + d_g8a = e4m3(hi/64), d_g8l = e4m3(lo*64), uint8 [G][C][Hp][sir_gallery_pitch8(Wp)]
 *             produced from d_ghi/d_glo by sir_gallery_pack_fp8c;
 *   templates: rows padded to 16 taps (Kpad = sir_template_kpad_fp8c), d_thi f16 + d_t8b = e4m3(hi/64),
 *             d_t8l = e4m3(lo*64), uint8 [C][ncols_alloc][Kpad], produced by sir_template_pack_fp8c.
 * Per 32-tap K stage the kernel issues two fp16 MMAs (hi*hi) and two e4m3 MMAs ((lo*64)(hi/64) and
 * (hi/64)(lo*64)) into the same fp32 accumulator. */
int sir_gallery_pitch8(int Wp);
int sir_gallery_pack_fp8c(const uint16_t* d_ghi, const uint16_t* d_glo, int G, int C, int Hp, int Wp,
                          uint8_t* d_g8a, uint8_t* d_g8l, void* stream);
int sir_template_kpad_fp8c(int Hm, int Wm);
int sir_template_pack_fp8c(const float* d_maps, int N, int C, int h, int w, int col0, int ncols_alloc,
                           uint16_t* d_thi, uint8_t* d_t8b, uint8_t* d_t8l, void* stream);
int sir_ncc_scores_fp8c(const uint16_t* d_ghi, const uint8_t* d_g8a, const uint8_t* d_g8l, const float* d_rnorm,
                        int G, int C, int Hp, int Wp, const uint16_t* d_thi, const uint8_t* d_t8b, const uint8_t* d_t8l,
                        int ncols, int ncols_alloc, int Hm, int Wm, const int32_t* d_col2probe, float* d_scores,
                        int score_ld, int g0, void* stream);

/* Multi-shape column tiles (ragged probe sets: every probe / scale variant its own template shape).
 * Templates of different true shapes are packed into the K layout of one BUCKET shape Hb x Wb, anchor on
 * anchor, zeros elsewhere (the numerator is unchanged); each chunk of sir_ncc_norm_chunk() (= 8) columns of the
 * block holds templates of one true shape and d_rnorm_tab[chunk] points at that shape's window-norm table
 * (similarity.py:57-65 depends on the TRUE template size).  d_rnorm_tab: device array of device pointers, 256 /
 * sir_ncc_norm_chunk() per 256-column tile, every entry valid (also those of the last tile's absent columns).
 * precision: SIR_PREC_FP16X3 (d_glo, d_tlo) or SIR_PREC_FP16_FP8C (e4m3 operands).
 * d_tile_rows (optional, NULL = none): int32 [tiles][2], per 256-column tile the rows [lo, hi) of the bucket's K layout that
 * hold a non-zero tap in ANY column of the tile (always including the anchor row Hb / 2); rows outside are skipped by
 * every role of the kernel exactly like the rows that only meet the "same" padding.  With the columns sorted by true
 * template height a tile of short templates costs what its own tallest template costs, not what the bucket's does. */
int sir_ncc_norm_chunk(void);
int sir_template_pack_embed(const float* d_maps, int N, int C, int h, int w, int Hb, int Wb, int col0, int ncols_alloc,
                            int precision, uint16_t* d_thi, uint16_t* d_tlo, uint8_t* d_t8b, uint8_t* d_t8l, void* stream);
int sir_ncc_scores_multi(const uint16_t* d_ghi, const uint16_t* d_glo, const uint8_t* d_g8a, const uint8_t* d_g8l,
                         const float* const* d_rnorm_tab, int G, int C, int Hp, int Wp,
                         const uint16_t* d_thi, const uint16_t* d_tlo, const uint8_t* d_t8b, const uint8_t* d_t8l,
                         int ncols, int ncols_alloc, int Hb, int Wb, const int32_t* d_col2probe, float* d_scores,
                         int score_ld, int g0, int precision, const int32_t* d_tile_rows, void* stream);

/* Screen + refine (SIR_PREC_FP16_REFINE).  The reference needs ONE number per (probe, gallery): the maximum of the
 * correlation surface over positions and variants (similarity.py:106-108, 365-367).  sir_ncc_screen computes the
 * whole surface on the tensor cores with plain fp16 operands (d_ghi, d_thi; good to ~2e-4 relative), max-reduces it
 * into d_approx exactly like sir_ncc_scores does into d_scores (caller zeroes it first), and leaves one 8-byte
 * record per (column n, gallery g, 16x8 position patch p) in d_rec[(n*G + g)*NP + p], NP = ceil(Hp/16)*ceil(Wp/8):
 *   .x = the patch maximum (float32 bits, score units), .y = up to three candidate rows (8 bits each: row>>3 = y
 *   offset, row&7 = x offset inside the patch) and, in the top byte, the number of rows within the margin
 *   tau(m) = tau_rel*|m| + tau_abs of the patch maximum (saturated; > 3 = "every position of the patch").
 * sir_ncc_refine then evaluates in float32 -- from d_g32 (sir_gallery_pack_f32) and d_t32p (sir_template_pack_screen),
 * the float32 values the fp16 operands were rounded from -- exactly the candidate positions of the records whose
 * maximum is within tau of the pair's screened maximum d_approx, and max-reduces the exact values into d_scores
 * (caller zeroes it first).  tau must cover twice the screening error.  d_rnorm (one template shape) or d_rnorm_tab
 * (multi-shape bucket, one table pointer per 16-column chunk as in sir_ncc_scores_multi): exactly one is non-NULL;
 * Hb x Wb is the K layout of the packed templates (= Hm x Wm for a single shape), rows padded to 8 taps.
 * variants_hint: about how many columns of the block belong to one probe (>= 1; sizes the refinement's tiles: with V
 * variants roughly one (column, gallery) cell in V has a candidate).  d_stats: NULL or 4 device counters ([0] positions evaluated, [1] records with more than 3 rows, [2] tiles with
 * work), accumulated.  sir_ncc_screen_rec_count: number of 8-byte records d_rec must hold.
 *
 * sir_gallery_pack_f32: sir_gallery_pack that also writes d_g32 [G][C][Hp][WP] float32 = (g - mean) * 2^e, rows
 * padded with zeros like d_ghi (NULL: not written).  sir_template_pack_screen: d_thi as sir_template_pack and
 * d_t32p [C][ncols_alloc][Kpad] float32 = (t - mean)/sqrt(E) * 2^10 in the same padded K layout; Hb x Wb >= the true
 * shape selects a bucket layout (anchor on anchor) as in sir_template_pack_embed. */
int sir_gallery_pack_f32(const float* d_gallery, int G, int C, int hg, int wg, uint16_t* d_ghi, uint16_t* d_glo, int32_t* d_gexp,
                         float* d_gz, float* d_g32, void* stream);
int sir_template_pack_screen(const float* d_maps, int N, int C, int h, int w, int Hb, int Wb, int col0, int ncols_alloc,
                             uint16_t* d_thi, float* d_t32p, const int32_t* d_gather, void* stream);
/* d_gather (NULL = none): index map [h*w] from sir_variant_index_map.  The maps handed to sir_template_pack_screen
 * are then the UNROTATED source maps (same number of cells); cell i of the h x w variant the columns show is source
 * cell d_gather[i] (-1 = 0, Pillow's fill).  sir_variant_index_map(h, w, angle, transpose, ...) writes the map of
 * Image.rotate(angle) on an h x w map (similarity.py:267), composed with a transposition when `transpose` != 0 (the
 * variant is then w x h).  The rotated / transposed variant maps never exist in HBM. */
int sir_variant_index_map(int h, int w, double angle, int transpose, int32_t* d_map, void* stream);
long long sir_ncc_screen_rec_count(int G, int Hp, int Wp, int ncols);
int sir_ncc_screen(const uint16_t* d_ghi, const float* d_rnorm, const float* const* d_rnorm_tab, int G, int C, int Hp, int Wp,
                   const uint16_t* d_thi, int ncols, int ncols_alloc, int Hb, int Wb, const int32_t* d_col2probe, float* d_approx,
                   int score_ld, int g0, float tau_rel, float tau_abs, void* d_rec, const int32_t* d_tile_rows, void* stream);
int sir_ncc_refine(const float* d_g32, const float* d_rnorm, const float* const* d_rnorm_tab, int G, int C, int Hp, int Wp,
                   const float* d_t32p, int ncols, int ncols_alloc, int Hb, int Wb, const int32_t* d_col2probe,
                   const float* d_approx, float* d_scores, int score_ld, int g0, float tau_rel, float tau_abs, const void* d_rec,
                   int variants_hint, unsigned long long* d_stats, void* stream);
/* cudaMemsetAsync(d_ptr, 0, bytes) on `stream`: the path zeroes its score / operand buffers through the library. */
int sir_memset_zero(void* d_ptr, size_t bytes, void* stream);

/* Test aid: one CTA per SM fills the SM's whole shared-memory carve-out with `byte`.  Shared memory is not cleared between
 * kernels; the parity tests run the correlation kernels after a 0xFF fill (NaN as fp16 / e4m3 / float32) to prove that no
 * MMA or reduction reads a cell the kernel itself has not written. */
int sir_debug_fill_shared_memory(int byte, void* stream);

/* Planning aid (host only, nothing is launched): estimated SM cycles per (gallery, 256-column tile, channel)
 * of the tensor-core kernel for this shape and precision mode -- the larger of the MMA time of the
 * non-skipped K stages and the shifted-entry generation time under the shared-memory plan the launch
 * would use.  The host uses it to choose map orientation and between FP16_FP8C and FP16X3. */
int sir_ncc_cost(int precision, int G, int Hp, int Wp, int Hm, int Wm, double* h_cost);

/* ------------------------------------------------------------------ ranking (K8, K9)
 * _get_rank (similarity.py:378-386) without the sort: d_true_score[q] is scores[q][true] on the
 * shard that owns it (else -inf; merged across shards by the caller with a max all-reduce),
 * d_count_gt[q] = #{g : scores[q][g] > true_score[q]}, d_count_ge likewise with >= (tie window,
 * excludes nothing), and the k best (score, global index = g0 + g) per probe, descending,
 * ties by lower index.  rank = 1 + sum over shards of count_gt. */
int sir_true_scores(const float* d_scores, int Q, int G, int score_ld, const int32_t* d_true_idx, int g0,
                    float* d_true_score, void* stream);
int sir_rank_topk(const float* d_scores, int Q, int G, int score_ld, const float* d_true_score, int g0, int k,
                  int32_t* d_count_gt, int32_t* d_count_ge, float* d_topk_val, int32_t* d_topk_idx, void* stream);
/* Merge P shards' [Q][k] lists (as gathered by ncclAllGather: [P][Q][k]) into the global k best. */
int sir_merge_topk(const float* d_vals, const int32_t* d_idx, int P, int Q, int k,
                   float* d_out_val, int32_t* d_out_idx, void* stream);

/* d_out[q][d_order[j]] = d_in[q][j] for q < Q, j < G: score columns computed shape group by shape group go back to
 * the caller's gallery order (d_order: int32 [G], a permutation). */
int sir_scatter_columns(const float* d_in, int Q, int G, int ld_in, const int32_t* d_order, float* d_out, int ld_out, void* stream);

/* ------------------------------------------------------------------ feature stage (K1-K3)
 * The truncated backbone of network.py:185-186,234-235, operator by operator.  Activations are
 * float32 NHWC [B][H][W][C] on the device.  Every producer maintains a running max |value| of its
 * output in a device float (`d_amax_out`, zeroed by the caller before the forward pass); the next
 * convolution reads it (`d_amax_in`) to pick the power-of-two scaling of its fp16 hi/lo operands.
 * act: 0 none, 1 SiLU, 2 ReLU.
 *
 * sir_feat_image_to_nhwc: ToTensor + (grayscale repeat) + Normalize, network.py:51-87.
 *   d_img uint8 [B][H][W] (in_ch 1) or [B][H][W][3]; h_mean/h_std: 3 host floats.
 * sir_feat_im2col_split: the operand pass of a convolution.  With a 1x1 kernel and Kp = C it is the split of a
 *   float32 NHWC tensor into the fp16 hi/lo planes sir_feat_conv reads; with kh*kw > 1 it gathers the explicit
 *   im2col matrix A [M = B*Ho*Wo][Kp], k = (ky*kw + kx)*C + c (used for strided convolutions, then multiplied by
 *   sir_feat_conv as a 1 x M "image" with C = Kp).  Values are scaled by 2^e(amax_in); d_chan_scale [B][C] or NULL
 *   multiplies the input per (image, channel) (squeeze-excitation scale).
 * sir_feat_dwconv: depthwise k x k Conv2d + folded BN bias + activation; weights [k][k][C].  If d_pool_part is
 *   not NULL it also receives the squeeze of a following SqueezeExcitation as partial sums over pixels,
 *   [B][parts][C] with parts = sir_feat_dwconv_pool_parts(k, stride, C, Ho, Wo) (fused into the 3x3 fast path).
 *   The 3x3 fast path can also (or only, d_out = NULL) write its result as operand planes for sir_feat_conv
 *   (d_out_hi/d_out_lo/d_exp_out, bound = amax_in * bound_mult + bound_add as there).
 * sir_feat_pool_sum: the same squeeze for any other producer, one part: [B][HW][C] -> [B][1][C].
 * sir_feat_se_scale: SqueezeExcitation._scale: avg = sum of parts / HW (d_avg [B][C] scratch), fc1 [S][C] + SiLU, fc2 given
 *   transposed as [S][C] + sigmoid -> d_scale [B][C].
 * sir_feat_maxpool: MaxPool2d (VGG).   sir_feat_nhwc_to_nchw: [B][HW][C] -> [B][C][HW]. */
/* sir_feat_clahe_to_nhwc: cv2.createCLAHE(clip_limit, (tiles_x, tiles_y)).apply (network.py:108-111,197-208) on
 * uint8 grayscale [B][H][W], bit exact, fused with ToTensor / repeat / Normalize.  d_lut: scratch of
 * B*tiles_x*tiles_y*256 bytes; d_clahe_u8 [B][H][W] or NULL receives the equalised uint8 image. */
int sir_feat_clahe_to_nhwc(const uint8_t* d_img, int B, int H, int W, double clip_limit, int tiles_x, int tiles_y,
                           const float* h_mean, const float* h_std, uint8_t* d_lut, uint8_t* d_clahe_u8, float* d_out,
                           float* d_amax_out, void* stream);
/* sir_feat_clahe_rgb_to_nhwc: the RGB branch of Model._clahe (network.py:199-204: cv2 RGB2LAB, CLAHE on L, LAB2RGB) on uint8
 * [B][H][W][3], bit exact, fused with ToTensor / Normalize.  d_rgb2lab / d_lab2rgb: the two 8-bit colour conversions as
 * 2^24-entry tables (uint32 c0 | c1 << 8 | c2 << 16 at index c0 << 16 | c1 << 8 | c2; the host builds them once with
 * cv2.cvtColor over all 2^24 pixels).  Scratch: d_l_plane [B][H][W] bytes, d_ab_plane [B][H][W] uint16, d_lut as above;
 * d_rgb_out [B][H][W][3] or NULL receives the equalised uint8 image. */
int sir_feat_clahe_rgb_to_nhwc(const uint8_t* d_rgb, int B, int H, int W, double clip_limit, int tiles_x, int tiles_y,
                               const float* h_mean, const float* h_std, const uint32_t* d_rgb2lab, const uint32_t* d_lab2rgb,
                               uint8_t* d_l_plane, uint16_t* d_ab_plane, uint8_t* d_lut, uint8_t* d_rgb_out, float* d_out,
                               float* d_amax_out, void* stream);
int sir_feat_image_to_nhwc(const uint8_t* d_img, int B, int H, int W, int in_ch, const float* h_mean, const float* h_std,
                           float* d_out, float* d_amax_out, void* stream);
int sir_feat_im2col_split(const float* d_in, const float* d_amax_in, int B, int H, int W, int C, int kh, int kw, int stride,
                          int pad, const float* d_chan_scale, int Kp, uint16_t* d_ahi, uint16_t* d_alo, void* stream);
/* sir_feat_conv: Conv2d (groups 1, square zero padding, any stride: strided patches are fetched with TMA element strides) + folded BN bias + activation (+ residual) as an
 * implicit GEMM: no im2col matrix.  d_xhi/d_xlo: the input split into fp16 hi/lo NHWC planes [B][H][W][C]
 * (sir_feat_im2col_split with a 1x1 kernel and Kp = C), C % 8 == 0.  Weights: d_wpack from sir_feat_conv_pack_weights,
 * packed for the tile sir_feat_conv_plan reports for this shape; K order k = (ky*kw + kx)*Cp + c, Cp = C rounded up to bk.  A GEMM over an explicit [M][K] matrix is the call with B=H=1, W=M, C=K, kh=kw=1.
 * Chaining without a split pass: with d_out_hi/d_out_lo the epilogue also (or only, d_out = NULL) writes the
 * result as the next convolution's operand planes [pixels][N] (N % 8 == 0), scaled by 2^e with
 * e = 15 - ceil(log2(amax_in*bound_mult + bound_add + amax_res)), an a-priori bound of |out| the caller derives
 * from the weights (max row L1 norm, max |bias|); e goes to *d_exp_out and is handed to the consumer as d_exp_in
 * (NULL = planes from sir_feat_im2col_split, scaled from the measured amax).
 * Replaces the torch Conv2d/BatchNorm/SiLU modules the reference runs at network.py:234-235. */
int sir_feat_conv_tile_n(int N);
/* The weights reach the kernel as ready-made shared-memory images, one contiguous block per pipeline stage (a plain bulk
 * copy; tensor-map loads of 32/64-byte weight rows are bound by the TMA unit's row rate).  sir_feat_conv_plan: tile width,
 * K granule (16 or 32 channels) and byte size of the packed weights for a shape; sir_feat_conv_pack_weights: [n_rows][Kp]
 * fp16 hi/lo matrices (k = (ky*kw + kx)*Cp + c) -> that layout, [n_tile][k_step][hi | lo][tile_n][granule] swizzled. */
int sir_feat_conv_plan(int B, int H, int W, int C, int kh, int kw, int pad, int stride, int bk, int N, int per_image,
                       int* tile_n, int* granule, long long* pack_bytes);
/* sir_feat_conv_scale_weights: one weight set per image with the SqueezeExcitation scale folded in,
 * W_b[n][k] = W[n][k] * d_scale[b][k mod Cp] (torchvision SqueezeExcitation.forward, scale * input, followed by the
 * 1x1 projection), packed like sir_feat_conv_pack_weights, B sets of pack_bytes; used with per_image = 1, which keeps
 * every row tile inside one image.  The projection then reads the depthwise output's own operand planes: no pass over
 * the activation between the depthwise convolution and the projection. */
int sir_feat_conv_scale_weights(const uint16_t* d_whi, const uint16_t* d_wlo, int n_rows, int Kp, int Cp, int C,
                                const float* d_scale, int B, int tile_n, int granule, uint8_t* d_pack, void* stream);
int sir_feat_conv_pack_weights(const uint16_t* d_whi, const uint16_t* d_wlo, int n_rows, int Kp, int tile_n, int granule,
                               uint8_t* d_pack, void* stream);
int sir_feat_conv(const uint16_t* d_xhi, const uint16_t* d_xlo, const float* d_amax_in, int B, int H, int W, int C, int kh,
                  int kw, int pad, int stride, int bk, const uint8_t* d_wpack, int pack_tile_n, int pack_granule, int per_image, int N, int w_exp,
                  const float* d_bias, const float* d_residual, int act, float* d_out, int ldc, float* d_amax_out,
                  const int32_t* d_exp_in, uint16_t* d_out_hi, uint16_t* d_out_lo, int32_t* d_exp_out, float bound_mult,
                  float bound_add, const float* d_amax_res, void* stream);
/* sir_feat_conv_c3k3: 3x3 Conv2d of a 3-channel NHWC image (the stem, K = 27) + bias + activation in float32 on the
 * CUDA cores; weights [27][Cout] with row = (ky*3 + kx)*3 + c, Cout % 8 == 0.  Optional operand-plane output as in
 * sir_feat_conv. */
int sir_feat_conv_c3k3(const float* d_in, const float* d_amax_in, int B, int H, int W, int stride, int pad, const float* d_w,
                       const float* d_bias, int Cout, int act, float* d_out, float* d_amax_out, uint16_t* d_out_hi,
                       uint16_t* d_out_lo, int32_t* d_exp_out, float bound_mult, float bound_add, void* stream);
int sir_feat_dwconv_pool_parts(int k, int stride, int C, int Ho, int Wo);
int sir_feat_dwconv(const float* d_in, int B, int H, int W, int C, int k, int stride, int pad, const float* d_w,
                    const float* d_bias, int act, float* d_out, float* d_amax_out, float* d_pool_part, const float* d_amax_in,
                    uint16_t* d_out_hi, uint16_t* d_out_lo, int32_t* d_exp_out, float bound_mult, float bound_add, void* stream);
int sir_feat_pool_sum(const float* d_in, int B, int HW, int C, float* d_pool_part, void* stream);
int sir_feat_se_scale(const float* d_pool_part, int B, int parts, int HW, int C, int S, const float* d_w1, const float* d_b1,
                      const float* d_w2t, const float* d_b2, float* d_avg, float* d_scale, void* stream);
int sir_feat_maxpool(const float* d_in, int B, int H, int W, int C, int k, int stride, int pad, float* d_out,
                     float* d_amax_out, void* stream);
/* per-channel y = act(x * scale[c] + shift[c]) (scale/shift both NULL: activation only) over `rows`
 * pixels of C channels; row strides ld_in / ld_out let it read or write a channel slice of a wider
 * NHWC buffer (DenseNet concatenation).  sir_feat_avgpool2d: AvgPool2d(k, stride), no padding. */
int sir_feat_affine_act(const float* d_in, long long rows, int C, int ld_in, int ld_out, const float* d_scale,
                        const float* d_shift, int act, float* d_out, float* d_amax_out, void* stream);
int sir_feat_avgpool2d(const float* d_in, int B, int H, int W, int C, int k, int stride, float* d_out, float* d_amax_out,
                       void* stream);
int sir_feat_nhwc_to_nchw(const float* d_in, int B, int HW, int C, float* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIR_H_ */
