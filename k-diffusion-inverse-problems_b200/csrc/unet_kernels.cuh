// Launchers of the non-GEMM UNet kernels (unet_kernels.cu, attention.cu), used by the UNet driver (unet.cu) and
// exported one by one through the C ABI for parity tests.  All activations bf16 NHWC, statistics fp32.
#pragma once
#include "kdip_common.cuh"

namespace kdip {

typedef __nv_bfloat16 bf16;

// resample codes shared by gn_apply / gn_bwd: the op applied AFTER the activation in the forward pass
enum { RS_NONE = 0, RS_AVGPOOL2 = 1, RS_NEAREST_UP2 = 2 };

// per-(image, channel) sum and sum of squares over pixels, accumulated into stats[N][C][2] (must be zeroed)
int launch_chan_stats(const bf16* x, int N, int P, int C, float* stats, cudaStream_t s);

// GroupNorm(32 groups, eps 1e-5) + optional FiLM folded into a per-(image, channel) affine u = A*x + B.
// stats0/stats1: per-channel sums of the (up to two, channel-concatenated) sources with C0 / C1 channels.
// film: [N][film_stride] fp32 with scale at film_off + c and shift at film_off + C + c, or NULL.
// Outputs: ab[N][C][2] = (A, B); mr[N][32][2] = (mean, rstd) (saved for the backward pass).
int launch_gn_finalize(const float* stats0, int C0, const float* stats1, int C1, int N, int P, const float* gamma,
                       const float* beta, const float* film, int film_stride, int film_off, float* ab, float* mr, cudaStream_t s);

// y = resample(act(A*x + B)) ; sources src0 [N,H,W,C0], src1 [N,H,W,C1] (or NULL), out [N,H',W',C0+C1]
// pool_out (RS_AVGPOOL2 only, optional): also writes avg_pool2d(x) of the raw input [N,H/2,W/2,C] (the Downsample skip path)
int launch_gn_apply(const bf16* src0, int C0, const bf16* src1, int C1, int N, int H, int W, const float* ab, int act_silu,
                    int resample, bf16* out, cudaStream_t s, bf16* pool_out = nullptr);

// backward, pass 1: red[N][C][2] += (sum_p g_u, sum_p g_u*x) with g_u = resample^T(g_y) * act'(A*x+B)
int launch_gn_bwd_reduce(const bf16* src0, int C0, const bf16* src1, int C1, int N, int H, int W, const float* ab, int act_silu,
                         int resample, const bf16* gy, float* red, cudaStream_t s);
// backward, finalize: k[N][C][4] = (k0, k1, k2, 0) so that g_x = k0*g_u + k1 + k2*x   (GroupNorm backward through mean and rstd)
int launch_gn_bwd_finalize(const float* red, const float* ab, const float* mr, const float* gamma, int N, int C, int P,
                           const float* film, int film_stride, int film_off, float* k, cudaStream_t s);
// backward, pass 2: g_x = k0*g_u + k1 + k2*x (+ extra) written to dst0 [N,H,W,C0] / dst1 [N,H,W,C1].
// extra_mode: 0 none, 1 tensor at x's resolution [N,H,W,C0+C1], 2 tensor at g_y's resolution (resample^T applied)
// skip_grad (single-source inputs only, optional): [N,H,W,C0] gradient that reached x through the skip stack, added as well
int launch_gn_bwd_apply(const bf16* src0, int C0, const bf16* src1, int C1, int N, int H, int W, const float* ab,
                        const float* k, int act_silu, int resample, const bf16* gy, const bf16* extra, int extra_mode,
                        bf16* dst0, bf16* dst1, cudaStream_t s, const bf16* skip_grad = nullptr);

// y[i] += alpha * x[i]  (fp32, tiny vectors such as biases)
int launch_axpy_f32(float* y, const float* x, float alpha, int n, cudaStream_t s);
// a += b (bf16, n elements, n % 8 == 0)
int launch_add_bf16(bf16* a, const bf16* b, size_t n, cudaStream_t s);

// Direct 3x3 conv (pad 1) for tiny input-channel counts: in fp32 NCHW [N,CIN,H,W] (scaled by in_scale[n] if non-null),
// w fp32 [9][CIN][Cout], bias [Cout] or NULL -> out bf16 NHWC [N,H,W,Cout].  CIN in {3, 6}, Cout % 8 == 0, Cout <= 256.
int launch_conv_small_cin(const float* in, const float* in_scale, const float* w, const float* bias, int N, int CIN, int H,
                          int W, int Cout, bf16* out, cudaStream_t s);
// fp32 OIHW [O][I][3][3] -> [9][CIN][Cout] fp32.  flip=0: CIN=I, Cout=O (forward).  flip=1: CIN=O, Cout=I, taps reversed
// (input-gradient of a conv whose OUTPUT has few channels, e.g. the UNet head).
int launch_pack_small(const float* w_oihw, int O, int I, int flip, float* dst, cudaStream_t s);

// im2col of a 3x3 / pad-1 neighbourhood for CIN <= 7 input channels: in fp32 NCHW (times in_scale[n]) -> out bf16 [N,H,W,64]
// with channel tap*CIN + ci (zero beyond 9*CIN and outside the image); and the matching GEMM weights [rows_pad][64].
int launch_im2col3x3(const float* in, const float* in_scale, int N, int CIN, int H, int W, bf16* out, cudaStream_t s);
int launch_pack_im2col_weight(const float* w_oihw, int O, int I, int flip, int rows_pad, bf16* dst, cudaStream_t s);

// 3x3 convs with <= 7 OUTPUT channels as a tap-folded 1x1 GEMM (P fp32 [N*H*W][ldp], column tap*CO + co) + this 9-neighbour gather
// into fp32 NCHW [N,CO,H,W] (+ bias[CO] or NULL); and the folded weights [dst_rows][cols] from the pack_weight layout.
int launch_tap_gather(const float* P, int ldp, const float* bias, int N, int CO, int H, int W, float* out, cudaStream_t s);
int launch_fold_taps(const bf16* src, int rows_pad, int CO, int cols, int dst_rows, bf16* dst, cudaStream_t s);

// timestep embedding (nn.py:103-121) + time_embed MLP (unet.py:473-477): semb[N][ted] = SiLU(W2 SiLU(W1 e(t) + b1) + b2)
int launch_time_embed(const float* t, int N, int mc, const float* w1, const float* b1, const float* w2, const float* b2,
                      float* semb, cudaStream_t s, bool dedupe = false);
// all ResBlocks' emb_layers Linear at once: out[N][R] = Wall[R][ted] . semb[n] + ball[R]   (unet.py:199-205,246)
// dedupe (time_embed) / t != nullptr (emb_proj): an image whose timestep equals an earlier image's re-uses that image's row - the
// two must be used together (time_embed then leaves the duplicate rows of semb unwritten)
int launch_emb_proj(const float* semb, const float* t, int N, int ted, const float* wall, const float* ball, int R, float* out, cudaStream_t s);

// QKVAttentionLegacy (unet.py:339-356).  qkv bf16 [N,T,3C], channel = head*3*ch + {q,k,v}*ch + c.
// out bf16 [N,T,C] (channel = head*ch + c); lse fp32 [N,heads,T] (log-sum-exp of the scaled scores, for the backward).
int launch_attention_fwd(const bf16* qkv, int N, int T, int heads, int ch, bf16* out, float* lse, cudaStream_t s);
// dqkv bf16 [N,T,3C] from d_out bf16 [N,T,C]; recomputes the probabilities from lse.
int launch_attention_bwd(const bf16* qkv, const bf16* out, const bf16* d_out, const float* lse, int N, int T, int heads, int ch,
                         bf16* dqkv, cudaStream_t s);

// tcgen05 / TMEM / TMA versions for T in {64, 128, 256} (attention_tc.cu); the launchers above dispatch to them when supported
bool attention_tc_supported(int T, int ch);
int launch_attention_fwd_tc(const bf16* qkv, int N, int T, int heads, bf16* out, float* lse, cudaStream_t s);
// streamed-block versions for T >= 384, T % 128 == 0 (attention_tcs.cu: the T = 1024 level of the ImageNet UNet)
bool attention_tcs_supported(int T, int ch);
int launch_attention_fwd_tcs(const bf16* qkv, int N, int T, int heads, bf16* out, float* lse, cudaStream_t s);
int launch_attention_bwd_tcs(const bf16* qkv, const bf16* out, const bf16* d_out, const float* lse, int N, int T, int heads,
                             bf16* dqkv, cudaStream_t s);
int launch_attention_bwd_tc(const bf16* qkv, const bf16* out, const bf16* d_out, const float* lse, int N, int T, int heads, bf16* dqkv,
                            cudaStream_t s);

}  // namespace kdip
