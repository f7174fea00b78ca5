// Internal launch functions shared between the per-kernel C ABI (rba_k_*) and the engine (engine.cu).
#pragma once
#include "common.cuh"

namespace rba {

int layernorm(const float* x, const float* gamma, const float* beta, int mode, int B, int H, int W, int C, int ws,
              int shift, float eps, float* y, uint16_t* y_hi, uint16_t* y_lo, cudaStream_t st);
int64_t groupnorm_ws_doubles(int B, int H, int W, int C, int groups);
int groupnorm(const float* x, int64_t x_bs, const float* gamma, const float* beta, int B, int H, int W, int C, int groups,
              float eps, const float* prev, int64_t prev_bs, int hp, int wp, int relu, float* y, uint16_t* y_hi,
              uint16_t* y_lo, int64_t y_bs, double* ws, cudaStream_t st);
int patch_embed(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean,
                const float* stdv, const float* conv_w, const float* conv_b, const float* gamma, const float* beta, int C,
                float* tokens, cudaStream_t st);
int stem_conv(const void* images, int img_dtype, int B, int H, int W, int Hp, int Wp, const float* mean, const float* stdv,
              const float* w, const float* bias, float* out, cudaStream_t st);
int maxpool3x3s2(const float* x, int B, int H, int W, int C, float* y, uint16_t* y_hi, uint16_t* y_lo, cudaStream_t st);
int bias_act_sub(const float* x, const float* bias, int B, int H, int W, int C, int stride, int relu, float* y, uint16_t* y_hi,
                 uint16_t* y_lo, cudaStream_t st);
int bn_fold_conv(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps, int O,
                 int64_t I, float* w_out, float* b_out, cudaStream_t st);
int ew_add(const float* a, const float* b, int64_t rows, int cols, int64_t period, float* y, uint16_t* y_hi,
           uint16_t* y_lo, cudaStream_t st);
int gemm(const rba_gemm_args& a, cudaStream_t st);
int conv3x3(const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi, const uint16_t* w_lo, int B, int H, int W,
            int Cin, int Cout, float* y, int backend, cudaStream_t st);
int window_attn(const float* qkv, const float* bias_table, int B, int H, int W, int C, int heads, int ws, int shift,
                uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st);
int window_attn_planes(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_table, const float* bias_prepared,
                       int B, int H, int W, int C, int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo,
                       cudaStream_t st);
int window_attn_tc(const uint16_t* qkv_hi, const uint16_t* qkv_lo, const float* bias_prepared, int B, int H, int W, int C,
                   int heads, int ws, int shift, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st, int tiled = 0);
int window_attn_bias_floats(int heads);
int window_attn_prepare_bias(const float* table, int heads, float* out, cudaStream_t st);
int mha(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, const uint8_t* mask, int B,
        int Lq, int Lk, int E, int heads, uint16_t* out_hi, uint16_t* out_lo, float* workspace, cudaStream_t st);
int64_t mha_workspace_floats(int B, int Lq, int Lk, int heads);
int attn_mask(const float* masks, int B, int Q, int h, int w, int th, int tw, uint8_t* out, cudaStream_t st);
int einsum_score_supported(int Q, int K, int D);
int einsum_score_launch(const uint16_t* e_hi, const uint16_t* e_lo, const float* bias, const uint16_t* y_hi,
                        const uint16_t* y_lo, const float* logits, int B, int Q, int K, int D, int h, int w, int H, int W,
                        int score_func, int include_void, float* rba, float* sem, cudaStream_t st);
int einsum_score2_supported(int Q, int K, int D);
int einsum_score2_launch(const uint16_t* e_hi, const uint16_t* e_lo, const float* bias, const uint16_t* y_hi, const uint16_t* y_lo,
                         const float* logits, int B, int Q, int K, int D, int h, int w, int H, int W, int include_void, float* rba,
                         cudaStream_t st);
int einsum_score3_supported(int Q, int K, int D);
int einsum_score3_launch(const uint16_t* e_hi, const uint16_t* e_lo, const float* bias, const uint16_t* y_hi, const uint16_t* y_lo,
                         const float* logits, int B, int Q, int K, int D, int h, int w, int H, int W, int include_void, float* rba,
                         cudaStream_t st);
int fused_score_variant();
int ood_pred_resize(const float* logits, int B, int h, int w, int H, int W, float* ood_pred, float* score, cudaStream_t st);
int msda_fused(const float* value, const int* Hs, const int* Ws, const float* oa, int B, int S, int M, int D, int L,
               int P, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st);

}  // namespace rba
