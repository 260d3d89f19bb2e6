// sir_ncc_scores: precision-mode dispatch for the correlation stage (K7).
#include "sir_common.cuh"

namespace sir {
int launch_ncc_simt(const float* d_gz, const float* d_rnorm, int G, int C, int Hp, int Wp, const float* d_t32, int ncols,
                    int ncols_alloc, int Hm, int Wm, const int32_t* d_col2probe, float* d_scores, int score_ld, int g0,
                    cudaStream_t st);
int launch_ncc_tc(const uint16_t* d_ghi, const uint16_t* d_glo, const uint8_t* d_g8a, const uint8_t* d_g8l, const float* d_rnorm, int G,
                  int C, int Hp, int Wp, const uint16_t* d_thi, const uint16_t* d_tlo, const uint8_t* d_t8b, const uint8_t* d_t8l,
                  int ncols, int ncols_alloc, int Hm, int Wm, const int32_t* d_col2probe, float* d_scores, int score_ld, int g0,
                  int passes, cudaStream_t st, double* cost_out, const float* const* d_rnorm_tab);
}  // namespace sir

using namespace sir;

extern "C" int sir_ncc_scores(const uint16_t* d_ghi, const uint16_t* d_glo, const int32_t* d_gexp, const float* d_gz,
                              const float* d_rnorm, int G, int C, int Hp, int Wp, const uint16_t* d_thi,
                              const uint16_t* d_tlo, const float* d_t32, int ncols, int ncols_alloc, int Hm, int Wm,
                              const int32_t* d_col2probe, float* d_scores, int score_ld, int g0, int precision,
                              void* stream) {
  (void)d_gexp;
  SIR_CHECK_ARG(d_rnorm && d_col2probe && d_scores, "sir_ncc_scores: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0, "sir_ncc_scores: empty gallery");
  SIR_CHECK_ARG(Hm > 0 && Wm > 0 && ncols > 0 && ncols <= ncols_alloc, "sir_ncc_scores: bad template block");
  SIR_CHECK_ARG(score_ld >= g0 + G, "sir_ncc_scores: score row (%d) shorter than g0+G (%d)", score_ld, g0 + G);
  cudaStream_t st = (cudaStream_t)stream;
  switch (precision) {
    case SIR_PREC_FP16X3:
      return launch_ncc_tc(d_ghi, d_glo, nullptr, nullptr, d_rnorm, G, C, Hp, Wp, d_thi, d_tlo, nullptr, nullptr, ncols, ncols_alloc,
                           Hm, Wm, d_col2probe, d_scores, score_ld, g0, 3, st, nullptr, nullptr);
    case SIR_PREC_FP16X1:
      return launch_ncc_tc(d_ghi, d_glo, nullptr, nullptr, d_rnorm, G, C, Hp, Wp, d_thi, d_tlo, nullptr, nullptr, ncols, ncols_alloc,
                           Hm, Wm, d_col2probe, d_scores, score_ld, g0, 1, st, nullptr, nullptr);
    case SIR_PREC_FP32_SIMT:
      return launch_ncc_simt(d_gz, d_rnorm, G, C, Hp, Wp, d_t32, ncols, ncols_alloc, Hm, Wm, d_col2probe, d_scores,
                             score_ld, g0, st);
    default:
      set_error("sir_ncc_scores: unknown precision mode %d", precision);
      return SIR_E_ARG;
  }
}

extern "C" int sir_ncc_scores_fp8c(const uint16_t* d_ghi, const uint8_t* d_g8a, const uint8_t* d_g8l, const float* d_rnorm, int G, int C,
                                   int Hp, int Wp, const uint16_t* d_thi, const uint8_t* d_t8b, const uint8_t* d_t8l, int ncols,
                                   int ncols_alloc, int Hm, int Wm, const int32_t* d_col2probe, float* d_scores, int score_ld, int g0,
                                   void* stream) {
  SIR_CHECK_ARG(d_rnorm && d_col2probe && d_scores, "sir_ncc_scores_fp8c: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0, "sir_ncc_scores_fp8c: empty gallery");
  SIR_CHECK_ARG(Hm > 0 && Wm > 0 && ncols > 0 && ncols <= ncols_alloc, "sir_ncc_scores_fp8c: bad template block");
  SIR_CHECK_ARG(score_ld >= g0 + G, "sir_ncc_scores_fp8c: score row (%d) shorter than g0+G (%d)", score_ld, g0 + G);
  return launch_ncc_tc(d_ghi, nullptr, d_g8a, d_g8l, d_rnorm, G, C, Hp, Wp, d_thi, nullptr, d_t8b, d_t8l, ncols, ncols_alloc, Hm, Wm,
                       d_col2probe, d_scores, score_ld, g0, 2, (cudaStream_t)stream, nullptr, nullptr);
}

// Multi-shape column tiles: the columns of the block were packed with sir_template_pack_embed into the K
// layout of a bucket shape Hb x Wb; every 16-column chunk holds templates of ONE true shape and
// d_rnorm_tab[chunk] (device array of device pointers, 16 per 256-column tile, every entry valid) is the
// window-norm table of that shape.  precision: SIR_PREC_FP16X3 or SIR_PREC_FP16_FP8C.
extern "C" int sir_ncc_scores_multi(const uint16_t* d_ghi, const uint16_t* d_glo, const uint8_t* d_g8a, const uint8_t* d_g8l,
                                    const float* const* d_rnorm_tab, int G, int C, int Hp, int Wp, const uint16_t* d_thi,
                                    const uint16_t* d_tlo, const uint8_t* d_t8b, const uint8_t* d_t8l, int ncols, int ncols_alloc, int Hb,
                                    int Wb, const int32_t* d_col2probe, float* d_scores, int score_ld, int g0, int precision,
                                    void* stream) {
  SIR_CHECK_ARG(d_rnorm_tab && d_col2probe && d_scores, "sir_ncc_scores_multi: null pointer");
  SIR_CHECK_ARG(G > 0 && C > 0 && Hp > 0 && Wp > 0, "sir_ncc_scores_multi: empty gallery");
  SIR_CHECK_ARG(Hb > 0 && Wb > 0 && ncols > 0 && ncols <= ncols_alloc, "sir_ncc_scores_multi: bad template block");
  SIR_CHECK_ARG(score_ld >= g0 + G, "sir_ncc_scores_multi: score row (%d) shorter than g0+G (%d)", score_ld, g0 + G);
  SIR_CHECK_ARG(precision == SIR_PREC_FP16X3 || precision == SIR_PREC_FP16_FP8C, "sir_ncc_scores_multi: precision %d not supported", precision);
  return launch_ncc_tc(d_ghi, d_glo, d_g8a, d_g8l, nullptr, G, C, Hp, Wp, d_thi, d_tlo, d_t8b, d_t8l, ncols, ncols_alloc, Hb, Wb, d_col2probe,
                       d_scores, score_ld, g0, precision == SIR_PREC_FP16X3 ? 3 : 2, (cudaStream_t)stream, nullptr, d_rnorm_tab);
}
