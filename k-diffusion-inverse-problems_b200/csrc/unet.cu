// ADM UNet driver: weight packing, workspace planning, forward and input-VJP as fixed launch lists.
//
// Mirrors guided_diffusion/unet.py:398-668 (UNetModel.__init__ / forward) with the hyper-parameters of
// condition/diffpir_utils/utils_model.py:353-387 (resblock_updown, scale-shift norm, learn_sigma, legacy attention
// order, head width 64).  State-dict names are the reference's (`input_blocks.3.0.in_layers.2.weight`, ...), so
// OpenAI checkpoints load unchanged (sample_condition_openai.py:130-132).
//
// Data layout in HBM (per batch of N images; everything lives in ONE caller-provided workspace):
//   activations   bf16 NHWC, one slot per block output / ResBlock mid tensor (kept for the VJP), two scratch slots for
//                 the normalised conv inputs;   statistics fp32 [N][C][2] per tensor;  GroupNorm affines fp32 [N][C][2].
//   weights       bf16 [taps*Cout][Cin] K-major (forward) and [taps*Cin][Cout] (input-gradient), packed once at create.
// The forward is: conv_in (direct) -> per block {stats -> GN finalize -> GN apply(+SiLU,+FiLM,+resample) -> tcgen05 conv}
// with the 1x1 skip conv folded into conv2 as extra K-blocks and identity/resampled skips added in the conv epilogue;
// torch.cat (unet.py:662) never materialises: consumers read the two sources.
// The VJP (autograd sites condition/condition.py:136,146,155,172,269; parameters never need gradients) walks the same
// plan backwards with dgrad convs (same kernel, flipped weights) and a 3-kernel GroupNorm backward.
#include <stdlib.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "unet_kernels.cuh"
#include "unet_plan.h"

namespace kdip {

// fp32 reference-precision engine (unet_fp32.cu), selected by kdip_unet_arch.precision
struct Fp32Engine;
int fp32_create(const kdip_unet_arch* arch, const std::map<std::string, std::pair<const float*, int64_t>>& src, Fp32Engine** out);
void fp32_destroy(Fp32Engine* e);
bool fp32_has_cov(const Fp32Engine* e);
int fp32_workspace_bytes(Fp32Engine* e, int N, size_t* bytes);
int fp32_prepare(Fp32Engine* e, int N, void* ws, size_t ws_bytes);
int fp32_forward(Fp32Engine* e, const float* x, const float* x_scale, const float* t, int N, float* out, float* cov_out, void* ws,
                 size_t ws_bytes, cudaStream_t s);
int fp32_vjp(Fp32Engine* e, const float* seed, int N, float* grad_x, void* ws, size_t ws_bytes, cudaStream_t s);
int fp32_feature(Fp32Engine* e, int N, float* feat, cudaStream_t s);

struct ConvPlan;  // conv_gemm.cu
int conv_plan_build(const kdip_conv_desc* d, ConvPlan* plan);
int conv_plan_launch(const ConvPlan* plan, cudaStream_t stream);
bool conv_can_fuse_gn_reduce(const kdip_conv_desc* d);
bool conv_can_fuse_gn_apply(const kdip_conv_desc* d);
ConvPlan* conv_plan_new();
void conv_plan_free(ConvPlan* p);

int pack_weight_ex(const float* w, int Cout, int Cin_total, int ci_off, int Cin_sub, int taps, int rows_pad, int cols_pad,
                   int flip, void* dst, cudaStream_t s);

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct ResWeights {
  const float *g1, *b1, *g2, *b2;       // GroupNorm affine
  bf16 *w1, *w1d, *w2, *w2d;            // conv weights, forward / dgrad
  bf16 *w1s0, *w1s1;                    // conv1 forward weights split by source (concatenated input read as two K-segments)
  bf16 *ws0, *ws1, *wsd;                // 1x1 skip: forward split by source, dgrad (all input channels)
  const float *bias1, *bias2;           // bias2 already includes the skip conv's bias
  int film_off;                         // offset of (scale, shift) in the emb_proj table
};
struct AttnWeights {
  const float *g, *b;
  bf16 *wqkv, *wqkvd, *wproj, *wprojd;
  const float *bqkv, *bproj;
};

struct Act {
  bf16* p;
  float* stats;
  int H, W, C;
};

// one step of a launch list; conv steps carry their algorithmic FLOPs for kdip_unet_profile
struct Op {
  std::function<int(cudaStream_t)> fn;
  bool is_conv = false;
  double flops = 0.0;
  template <class F>
  Op(F f) : fn(f) {}
  int operator()(cudaStream_t s) const { return fn(s); }
};

static double conv_flops(const kdip_conv_desc& d) {
  double k = 0;
  for (int i = 0; i < d.nseg; ++i) k += (double)d.seg[i].C * d.seg[i].taps;
  return 2.0 * d.N * d.H * d.W * (double)d.Cout * k;
}

}  // namespace kdip

using namespace kdip;

struct kdip_unet {
  kdip::Fp32Engine* f32 = nullptr;               // set instead of everything below when arch.precision == KDIP_PRECISION_FP32
  kdip_unet_arch arch;
  std::vector<BlockDesc> plan;
  std::map<std::string, const float*> fp32;   // small fp32 params resident on device (owned, see owned[])
  std::vector<void*> owned;                    // cudaMalloc'ed by create
  std::vector<ResWeights> resw;                // per plan entry (unused entries for other kinds)
  std::vector<AttnWeights> attw;
  bf16 *w_in_i2c = nullptr, *w_headd_i2c = nullptr;        // im2col GEMM weights [C0][64] (first conv, head input-gradient)
  bf16 *w_ind = nullptr, *w_head = nullptr, *w_cov = nullptr;
  bf16 *w_head_fold = nullptr, *w_ind_fold = nullptr;       // tap-folded [64][C0] / [32][C0] (unet_kernels.cu::tap_gather)
  const float *b_in = nullptr, *b_head = nullptr, *b_cov = nullptr, *g_head = nullptr, *be_head = nullptr;
  float *wall = nullptr, *ball = nullptr;      // concatenated emb_layers
  int R = 0;                                   // rows of the emb_proj table
  bool has_cov = false;

  // ---- per-(N, workspace) launch plan ----
  int planned_N = 0;
  void* planned_ws = nullptr;
  size_t planned_bytes = 0;
  std::vector<Op> fwd_ops, bwd_ops;
  std::vector<ConvPlan*> conv_plans;
  void* stats_base = nullptr;
  size_t stats_bytes = 0;
  void* red_base = nullptr;
  size_t red_bytes = 0;
  // per-call I/O (read by the ops through `this`)
  const float* io_x = nullptr;
  const float* io_xscale = nullptr;
  const float* io_t = nullptr;
  float* io_out = nullptr;
  float* io_cov = nullptr;
  const float* io_seed = nullptr;
  float* io_grad = nullptr;
  // convs whose output is a caller buffer (address known per call): rebuilt when the pointer changes
  kdip::ConvPlan *head_plan = nullptr, *cov_plan = nullptr, *ind_plan = nullptr;
  float *head_out = nullptr, *cov_out_ptr = nullptr, *ind_out = nullptr;
  const void* hlast_ptr = nullptr;   // pre-head feature (bf16 NHWC)
  const void* scrA_ptr = nullptr;    // normalised head input
  void* scrB_ptr = nullptr;          // free scratch at head time: tap-folded head GEMM output P (fp32 [N*S*S][64])
  void* g3_ptr = nullptr;            // free scratch at the end of the VJP: P of the first layer's input-gradient (fp32 [N*S*S][32])
  const void* gin_final = nullptr;   // gradient wrt the first conv's output
};

namespace kdip {

// ---------------------------------------------------------------------------------------------------------------------
// block plan (same walk as UNetModel.__init__, unet.py:482-618)
// ---------------------------------------------------------------------------------------------------------------------
static bool in_list(const int* a, int n, int v) {
  for (int i = 0; i < n; ++i)
    if (a[i] == v) return true;
  return false;
}

void build_block_plan(const kdip_unet_arch& a, std::vector<BlockDesc>& plan) {
  plan.clear();
  const int mc = a.model_channels;
  auto push = [&](const std::string& prefix, int kind, int cin, int cout, int updown, int stage, int block, int skip_ch) {
    BlockDesc b;
    b.prefix = prefix; b.kind = kind; b.cin = cin; b.cout = cout; b.updown = updown; b.stage = stage; b.block = block;
    b.skip_ch = skip_ch; b.first_of_block = b.last_of_block = false;
    plan.push_back(b);
  };
  int ch = (int)(a.channel_mult[0] * mc);
  push("input_blocks.0.0", 0, a.in_channels, ch, 0, 0, 0, 0);
  std::vector<int> chans{ch};
  int ds = 1, bi = 1;
  for (int level = 0; level < a.n_mult; ++level) {
    for (int r = 0; r < a.num_res_blocks; ++r) {
      const int cout = (int)(a.channel_mult[level] * mc);
      push("input_blocks." + std::to_string(bi) + ".0", 1, ch, cout, 0, 0, bi, 0);
      ch = cout;
      if (in_list(a.attention_ds, a.n_att, ds)) push("input_blocks." + std::to_string(bi) + ".1", 2, ch, ch, 0, 0, bi, 0);
      chans.push_back(ch);
      ++bi;
    }
    if (level != a.n_mult - 1) {
      push("input_blocks." + std::to_string(bi) + ".0", 1, ch, ch, 1, 0, bi, 0);
      chans.push_back(ch);
      ++bi;
      ds *= 2;
    }
  }
  push("middle_block.0", 1, ch, ch, 0, 1, 0, 0);
  push("middle_block.1", 2, ch, ch, 0, 1, 0, 0);
  push("middle_block.2", 1, ch, ch, 0, 1, 0, 0);
  int bo = 0;
  for (int level = a.n_mult - 1; level >= 0; --level) {
    for (int i = 0; i <= a.num_res_blocks; ++i) {
      const int ich = chans.back();
      chans.pop_back();
      const int cout = (int)(mc * a.channel_mult[level]);
      int li = 0;
      push("output_blocks." + std::to_string(bo) + "." + std::to_string(li++), 1, ch + ich, cout, 0, 2, bo, ich);
      ch = cout;
      if (in_list(a.attention_ds, a.n_att, ds)) push("output_blocks." + std::to_string(bo) + "." + std::to_string(li++), 2, ch, ch, 0, 2, bo, 0);
      if (level && i == a.num_res_blocks) {
        push("output_blocks." + std::to_string(bo) + "." + std::to_string(li++), 1, ch, ch, 2, 2, bo, 0);
        ds /= 2;
      }
      ++bo;
    }
  }
  for (size_t i = 0; i < plan.size(); ++i) {
    plan[i].first_of_block = (i == 0) || plan[i - 1].stage != plan[i].stage || plan[i - 1].block != plan[i].block;
    plan[i].last_of_block = (i + 1 == plan.size()) || plan[i + 1].stage != plan[i].stage || plan[i + 1].block != plan[i].block;
  }
}

static int pad_rows(int c) { return c <= 16 ? 16 : (c <= 32 ? 32 : ((c + 63) / 64) * 64); }

}  // namespace kdip

// ---------------------------------------------------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------------------------------------------------
static int dev_alloc(kdip_unet* u, size_t bytes, void** out) {
  void* p = nullptr;
  KDIP_CUDA(cudaMalloc(&p, bytes));
  u->owned.push_back(p);
  *out = p;
  return KDIP_OK;
}

extern "C" void kdip_unet_destroy(kdip_unet* u) {
  if (!u) return;
  if (u->f32) fp32_destroy(u->f32);
  for (void* p : u->owned) cudaFree(p);
  for (ConvPlan* c : u->conv_plans) conv_plan_free(c);
  delete u;
}

extern "C" int kdip_unet_create(const kdip_unet_arch* arch, int n_tensors, const char* const* names, const float* const* ptrs,
                                const int64_t* numels, kdip_unet** out) {
  KDIP_REQUIRE(arch && names && ptrs && numels && out, KDIP_EINVAL, "unet_create: null argument");
  KDIP_REQUIRE(arch->n_mult >= 1 && arch->n_mult <= 8 && arch->n_att >= 0 && arch->n_att <= 8, KDIP_EINVAL, "unet_create: bad arch");
  KDIP_REQUIRE(arch->num_head_channels == 64, KDIP_ESHAPE, "unet_create: num_head_channels must be 64 (got %d)", arch->num_head_channels);
  KDIP_REQUIRE(arch->in_channels == 3 && arch->out_channels == 6, KDIP_ESHAPE, "unet_create: in/out channels must be 3/6");
  KDIP_REQUIRE(arch->model_channels % 64 == 0, KDIP_ESHAPE, "unet_create: model_channels must be a multiple of 64");
  KDIP_REQUIRE(arch->precision == KDIP_PRECISION_BF16 || arch->precision == KDIP_PRECISION_FP32, KDIP_EINVAL,
               "unet_create: precision must be KDIP_PRECISION_BF16 or KDIP_PRECISION_FP32 (got %d)", arch->precision);
  int rc = kdip_device_check(nullptr);
  if (rc != KDIP_OK) return rc;
  kdip_unet* u = new kdip_unet();
  u->arch = *arch;
  build_block_plan(*arch, u->plan);
  std::map<std::string, std::pair<const float*, int64_t>> src;
  for (int i = 0; i < n_tensors; ++i) src[names[i]] = std::make_pair(ptrs[i], numels[i]);
  if (arch->precision == KDIP_PRECISION_FP32) {
    rc = fp32_create(arch, src, &u->f32);
    if (rc != KDIP_OK) { delete u; return rc; }
    u->has_cov = fp32_has_cov(u->f32);
    *out = u;
    return KDIP_OK;
  }
  cudaStream_t s = 0;
  auto fail = [&](int code) { kdip_unet_destroy(u); return code; };

  // fetch a source tensor, checking its element count
  auto get = [&](const std::string& name, int64_t numel, const float** p) -> int {
    auto it = src.find(name);
    KDIP_REQUIRE(it != src.end(), KDIP_EINVAL, "unet_create: missing tensor '%s'", name.c_str());
    KDIP_REQUIRE(it->second.second == numel, KDIP_ESHAPE, "unet_create: tensor '%s' has %lld elements, expected %lld", name.c_str(),
                 (long long)it->second.second, (long long)numel);
    *p = it->second.first;
    return KDIP_OK;
  };
  // device-resident fp32 copy owned by the handle
  auto keep = [&](const std::string& name, int64_t numel, const float** dst) -> int {
    const float* p;
    int r = get(name, numel, &p);
    if (r) return r;
    void* d;
    r = dev_alloc(u, (size_t)numel * 4, &d);
    if (r) return r;
    KDIP_CUDA(cudaMemcpyAsync(d, p, (size_t)numel * 4, cudaMemcpyDeviceToDevice, s));
    *dst = (const float*)d;
    return KDIP_OK;
  };
  auto pack = [&](const std::string& name, int Cout, int Cin_total, int ci_off, int Cin_sub, int taps, int flip, bf16** dst) -> int {
    const float* p;
    int r = get(name, (int64_t)Cout * Cin_total * taps, &p);
    if (r) return r;
    const int rows = flip ? Cin_sub : Cout, cols = flip ? Cout : Cin_sub;
    const int rp = pad_rows(rows), cp = ((cols + 63) / 64) * 64;
    void* d;
    r = dev_alloc(u, (size_t)taps * rp * cp * 2, &d);
    if (r) return r;
    *dst = (bf16*)d;
    return pack_weight_ex(p, Cout, Cin_total, ci_off, Cin_sub, taps, rp, cp, flip, d, s);
  };
#define TRY(x) do { int _r = (x); if (_r != KDIP_OK) return fail(_r); } while (0)

  const int mc = arch->model_channels, ted = 4 * mc;
  const float *tw1, *tb1, *tw2, *tb2;
  TRY(keep("time_embed.0.weight", (int64_t)ted * mc, &tw1));
  TRY(keep("time_embed.0.bias", ted, &tb1));
  TRY(keep("time_embed.2.weight", (int64_t)ted * ted, &tw2));
  TRY(keep("time_embed.2.bias", ted, &tb2));
  u->fp32["tw1"] = tw1; u->fp32["tb1"] = tb1; u->fp32["tw2"] = tw2; u->fp32["tb2"] = tb2;

  // emb_layers table
  int R = 0;
  for (auto& b : u->plan) if (b.kind == 1) R += 2 * b.cout;
  u->R = R;
  TRY(dev_alloc(u, (size_t)R * ted * 4, (void**)&u->wall));
  TRY(dev_alloc(u, (size_t)R * 4, (void**)&u->ball));

  u->resw.resize(u->plan.size());
  u->attw.resize(u->plan.size());
  int film_off = 0;
  for (size_t i = 0; i < u->plan.size(); ++i) {
    const BlockDesc& b = u->plan[i];
    const std::string& p = b.prefix;
    if (b.kind == 0) {
      const float* w;
      TRY(get(p + ".weight", (int64_t)b.cout * 3 * 9, &w));
      if (b.cout % 64 != 0) { set_error("unet_create: first conv width %d must be a multiple of 64", b.cout); return fail(KDIP_ESHAPE); }
      TRY(dev_alloc(u, (size_t)b.cout * 64 * 2, (void**)&u->w_in_i2c));
      TRY(launch_pack_im2col_weight(w, b.cout, 3, 0, b.cout, u->w_in_i2c, s));
      TRY(keep(p + ".bias", b.cout, &u->b_in));
      TRY(pack(p + ".weight", b.cout, 3, 0, 3, 9, 1, &u->w_ind));   // dgrad: rows = 3 (pad 16), cols = cout
      TRY(dev_alloc(u, (size_t)32 * b.cout * 2, (void**)&u->w_ind_fold));
      TRY(launch_fold_taps(u->w_ind, 16, 3, b.cout, 32, u->w_ind_fold, s));
    } else if (b.kind == 1) {
      ResWeights& rw = u->resw[i];
      memset(&rw, 0, sizeof(rw));
      const int ci = b.cin, co = b.cout;
      TRY(keep(p + ".in_layers.0.weight", ci, &rw.g1));
      TRY(keep(p + ".in_layers.0.bias", ci, &rw.b1));
      TRY(keep(p + ".out_layers.0.weight", co, &rw.g2));
      TRY(keep(p + ".out_layers.0.bias", co, &rw.b2));
      TRY(pack(p + ".in_layers.2.weight", co, ci, 0, ci, 9, 0, &rw.w1));
      TRY(pack(p + ".in_layers.2.weight", co, ci, 0, ci, 9, 1, &rw.w1d));
      if (b.skip_ch > 0 && (ci - b.skip_ch) % 64 == 0 && b.skip_ch % 64 == 0) {
        TRY(pack(p + ".in_layers.2.weight", co, ci, 0, ci - b.skip_ch, 9, 0, &rw.w1s0));
        TRY(pack(p + ".in_layers.2.weight", co, ci, ci - b.skip_ch, b.skip_ch, 9, 0, &rw.w1s1));
      }
      TRY(pack(p + ".out_layers.3.weight", co, co, 0, co, 9, 0, &rw.w2));
      TRY(pack(p + ".out_layers.3.weight", co, co, 0, co, 9, 1, &rw.w2d));
      TRY(keep(p + ".in_layers.2.bias", co, &rw.bias1));
      const float* b2;
      TRY(get(p + ".out_layers.3.bias", co, &b2));
      float* b2sum;
      TRY(dev_alloc(u, (size_t)co * 4, (void**)&b2sum));
      KDIP_CUDA(cudaMemcpyAsync(b2sum, b2, (size_t)co * 4, cudaMemcpyDeviceToDevice, s));
      if (ci != co) {
        const int c0 = ci - b.skip_ch, c1 = b.skip_ch;
        TRY(pack(p + ".skip_connection.weight", co, ci, 0, c0, 1, 0, &rw.ws0));
        if (c1 > 0) TRY(pack(p + ".skip_connection.weight", co, ci, c0, c1, 1, 0, &rw.ws1));
        TRY(pack(p + ".skip_connection.weight", co, ci, 0, ci, 1, 1, &rw.wsd));
        const float* bs;
        TRY(get(p + ".skip_connection.bias", co, &bs));
        TRY(launch_axpy_f32(b2sum, bs, 1.f, co, s));
      } else {
        KDIP_REQUIRE(b.skip_ch == 0, KDIP_ESHAPE, "unet_create: identity skip over a concatenated input is unsupported (%s)", p.c_str());
      }
      rw.bias2 = b2sum;
      rw.film_off = film_off;
      const float *ew, *eb;
      TRY(get(p + ".emb_layers.1.weight", (int64_t)2 * co * ted, &ew));
      TRY(get(p + ".emb_layers.1.bias", 2 * co, &eb));
      KDIP_CUDA(cudaMemcpyAsync(u->wall + (size_t)film_off * ted, ew, (size_t)2 * co * ted * 4, cudaMemcpyDeviceToDevice, s));
      KDIP_CUDA(cudaMemcpyAsync(u->ball + film_off, eb, (size_t)2 * co * 4, cudaMemcpyDeviceToDevice, s));
      film_off += 2 * co;
    } else {
      AttnWeights& aw = u->attw[i];
      memset(&aw, 0, sizeof(aw));
      const int c = b.cin;
      TRY(keep(p + ".norm.weight", c, &aw.g));
      TRY(keep(p + ".norm.bias", c, &aw.b));
      TRY(pack(p + ".qkv.weight", 3 * c, c, 0, c, 1, 0, &aw.wqkv));
      TRY(pack(p + ".qkv.weight", 3 * c, c, 0, c, 1, 1, &aw.wqkvd));
      TRY(pack(p + ".proj_out.weight", c, c, 0, c, 1, 0, &aw.wproj));
      TRY(pack(p + ".proj_out.weight", c, c, 0, c, 1, 1, &aw.wprojd));
      TRY(keep(p + ".qkv.bias", 3 * c, &aw.bqkv));
      TRY(keep(p + ".proj_out.bias", c, &aw.bproj));
    }
  }
  {
    const int c0 = (int)(arch->channel_mult[0] * mc);
    TRY(keep("out.0.weight", c0, &u->g_head));
    TRY(keep("out.0.bias", c0, &u->be_head));
    TRY(pack("out.2.weight", 6, c0, 0, c0, 9, 0, &u->w_head));      // rows 6 -> 16
    TRY(dev_alloc(u, (size_t)64 * c0 * 2, (void**)&u->w_head_fold));
    TRY(launch_fold_taps(u->w_head, 16, 6, c0, 64, u->w_head_fold, s));
    const float* hb;
    TRY(get("out.2.bias", 6, &hb));
    float* hb16;
    TRY(dev_alloc(u, 16 * 4, (void**)&hb16));
    KDIP_CUDA(cudaMemsetAsync(hb16, 0, 16 * 4, s));
    KDIP_CUDA(cudaMemcpyAsync(hb16, hb, 6 * 4, cudaMemcpyDeviceToDevice, s));
    u->b_head = hb16;
    const float* hw;
    TRY(get("out.2.weight", (int64_t)6 * c0 * 9, &hw));
    TRY(dev_alloc(u, (size_t)c0 * 64 * 2, (void**)&u->w_headd_i2c));
    TRY(launch_pack_im2col_weight(hw, 6, c0, 1, c0, u->w_headd_i2c, s));
    // optional DWT-Var covariance head (k_diffusion/external.py:141): Conv2d(C, 6, 1) on the pre-head feature
    if (src.count("out_cov.weight")) {
      TRY(pack("out_cov.weight", 6, c0, 0, c0, 1, 0, &u->w_cov));
      const float* cb;
      TRY(get("out_cov.bias", 6, &cb));
      float* cb16;
      TRY(dev_alloc(u, 16 * 4, (void**)&cb16));
      KDIP_CUDA(cudaMemsetAsync(cb16, 0, 16 * 4, s));
      KDIP_CUDA(cudaMemcpyAsync(cb16, cb, 6 * 4, cudaMemcpyDeviceToDevice, s));
      u->b_cov = cb16;
      u->has_cov = true;
    }
  }
  KDIP_CUDA(cudaStreamSynchronize(s));
#undef TRY
  *out = u;
  return KDIP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// launch plan for a batch size and workspace
// ---------------------------------------------------------------------------------------------------------------------
namespace kdip {

struct Builder {
  kdip_unet* u;
  int N;
  char* base;      // nullptr in the sizing pass
  size_t cur = 0;
  bool emit;
  int rc = KDIP_OK;

  void* alloc(size_t bytes) {
    size_t o = cur;
    cur += (bytes + 255) & ~(size_t)255;
    return base ? (void*)(base + o) : (void*)nullptr;
  }
  Act new_act(int H, int W, int C) {
    Act a;
    a.H = H; a.W = W; a.C = C;
    a.p = (bf16*)alloc((size_t)N * H * W * C * 2);
    a.stats = nullptr;
    return a;
  }
};

static ConvPlan* make_conv(kdip_unet* u, const kdip_conv_desc& d, int* rc) {
  ConvPlan* p = conv_plan_new();
  *rc = conv_plan_build(&d, p);
  if (*rc != KDIP_OK) {
    conv_plan_free(p);
    return nullptr;
  }
  u->conv_plans.push_back(p);
  return p;
}

}  // namespace kdip

static int build_launch_plan(kdip_unet* u, int N, void* ws, size_t ws_bytes, size_t* need_bytes) {
  const bool emit = (ws != nullptr);
  const kdip_unet_arch& a = u->arch;
  const int mc = a.model_channels, ted = 4 * mc, S = a.image_size;
  if (emit) {
    for (ConvPlan* c : u->conv_plans) conv_plan_free(c);
    u->conv_plans.clear();
    u->fwd_ops.clear();
    u->bwd_ops.clear();
  }
  Builder B{u, N, (char*)ws, 0, emit};
  std::vector<Op>& F = u->fwd_ops;
  std::vector<Op> Bk;   // backward ops collected in forward order of creation, reversed per block at the end
  int rc = KDIP_OK;

  // ---- zero-initialised accumulator regions: channel statistics (forward) and GN-backward reductions -------------
  // sized by a first walk over the plan
  size_t stats_floats = 0, red_floats = 0;
  {
    int H = S;
    auto add_stats = [&](int C) { stats_floats += (size_t)N * C * 2; };
    for (auto& b : u->plan) {
      if (b.kind == 0) add_stats(b.cout);
      else if (b.kind == 1) {
        if (b.updown == 1) H /= 2; else if (b.updown == 2) H *= 2;
        add_stats(b.cout); add_stats(b.cout);                 // h1, out
        red_floats += (size_t)N * (b.cin + b.cout) * 2;
      } else { add_stats(b.cout); red_floats += (size_t)N * b.cin * 2; }
    }
    red_floats += (size_t)N * (size_t)(a.channel_mult[0] * mc) * 2;
    (void)H;
  }
  float* stats_region = (float*)B.alloc(stats_floats * 4);
  float* red_region = (float*)B.alloc(red_floats * 4);
  size_t stats_used = 0, red_used = 0;
  auto take_stats = [&](int C) { float* p = stats_region ? stats_region + stats_used : nullptr; stats_used += (size_t)N * C * 2; return p; };
  auto take_red = [&](int C) { float* p = red_region ? red_region + red_used : nullptr; red_used += (size_t)N * C * 2; return p; };
  if (emit) { u->stats_base = stats_region; u->stats_bytes = stats_floats * 4; u->red_base = red_region; u->red_bytes = red_floats * 4; }

  // ---- small fp32 buffers -----------------------------------------------------------------------------------------
  float* semb = (float*)B.alloc((size_t)N * ted * 4);
  float* film = (float*)B.alloc((size_t)N * u->R * 4);
  int maxC = 0;
  for (auto& b : u->plan) maxC = std::max(maxC, std::max(b.cin, b.cout));
  float* kbuf = (float*)B.alloc((size_t)N * maxC * 4 * 4);     // GN-backward coefficients (transient)

  // scratch activations: the largest tensor any step produces
  size_t max_elems = 0;
  {
    int H = S;
    for (auto& b : u->plan) {
      if (b.kind == 1) {
        int Hin = H;
        if (b.updown == 1) H /= 2; else if (b.updown == 2) H *= 2;
        int Hm = std::max(H, Hin);
        max_elems = std::max(max_elems, (size_t)Hm * Hm * std::max(b.cin, b.cout));
      } else if (b.kind == 2) {
        max_elems = std::max(max_elems, (size_t)H * H * 3 * b.cin);
      } else max_elems = std::max(max_elems, (size_t)H * H * b.cout);
    }
  }
  const size_t scratch_bytes = (size_t)N * max_elems * 2;
  bf16* scrA = (bf16*)B.alloc(scratch_bytes);   // a1 / a (normalised conv inputs)
  bf16* scrB = (bf16*)B.alloc(scratch_bytes);   // a2
  bf16* g0 = (bf16*)B.alloc(scratch_bytes);     // backward: gradient ping-pong and transients
  bf16* g1 = (bf16*)B.alloc(scratch_bytes);
  bf16* g2 = (bf16*)B.alloc(scratch_bytes);
  bf16* g3 = (bf16*)B.alloc(scratch_bytes);
  bf16* g4 = (bf16*)B.alloc(scratch_bytes);

  double last_conv_flops = 0.0;
  auto conv = [&](const kdip_conv_desc& d) -> ConvPlan* {
    if (!emit) return nullptr;
    int r;
    ConvPlan* p = make_conv(u, d, &r);
    if (r != KDIP_OK && rc == KDIP_OK) rc = r;
    last_conv_flops = conv_flops(d);
    return p;
  };
  auto add_conv_op = [&](std::vector<Op>& ops, ConvPlan* p) {
    if (emit && p) {
      Op o([p](cudaStream_t s) { return conv_plan_launch(p, s); });
      o.is_conv = true;
      o.flops = last_conv_flops;
      ops.push_back(o);
    }
  };
  // GroupNorm statistics of every block output ride in the producing conv's epilogue (kdip_conv_desc.chan_stats);
  // KDIP_UNFUSED_STATS=1 restores the separate chan_stats pass (A/B measurements)
  const bool fused_stats = getenv("KDIP_UNFUSED_STATS") == nullptr;
  auto stats_op = [&](std::vector<Op>& ops, const Act& t) {
    if (!emit || fused_stats) return;
    const bf16* p = t.p; float* st = t.stats; int P = t.H * t.W, C = t.C, n = N;
    ops.push_back([=](cudaStream_t s) { return launch_chan_stats(p, n, P, C, st, s); });
  };

  // ================================= forward =================================
  if (emit) {
    kdip_unet* uu = u;
    const float *tw1 = u->fp32["tw1"], *tb1 = u->fp32["tb1"], *tw2 = u->fp32["tw2"], *tb2 = u->fp32["tb2"];
    int n = N, R = u->R;
    F.push_back([=](cudaStream_t s) {
      KDIP_CUDA(cudaMemsetAsync(uu->stats_base, 0, uu->stats_bytes, s));
      return launch_time_embed(uu->io_t, n, mc, tw1, tb1, tw2, tb2, semb, s, true);
    });
    F.push_back([=](cudaStream_t s) { return launch_emb_proj(semb, uu->io_t, n, ted, uu->wall, uu->ball, R, film, s); });
  }

  struct SavedRes { Act src0, src1, h1, out; float *ab1, *mr1, *ab2, *mr2; int Hin, Win, Ho, Wo; bool two; };
  struct SavedAttn { Act x, qkv, att, out; float *ab, *mr, *lse; };
  std::vector<Act> hs;          // skip stack (forward)
  std::vector<bf16*> hs_grad;   // gradient slot of each pushed tensor
  struct Pending { bf16* gslot; };
  Act h;
  int H = S;
  // per-block backward op lists, executed in reverse block order
  std::vector<std::vector<Op>> bwd_blocks;
  // gradient bookkeeping resolved at build time: `gcur` = buffer holding d/d(h) when walking backwards
  struct BwdCtx { int dummy; };

  // We build backward ops while walking forward, but they need to know which buffer holds the incoming gradient.  The
  // walk backwards is deterministic, so we first record per-block descriptors, then emit backward ops in a second loop.
  struct Rec { int kind; size_t idx; SavedRes r; SavedAttn t; Act in0; bool pushes; int push_id; int pop_id; };
  std::vector<Rec> recs;

  for (size_t i = 0; i < u->plan.size(); ++i) {
    const BlockDesc& b = u->plan[i];
    Rec rec;
    rec.kind = b.kind; rec.idx = i; rec.pushes = false; rec.push_id = -1; rec.pop_id = -1;
    if (b.kind == 0) {
      h = B.new_act(H, H, b.cout);
      h.stats = take_stats(b.cout);
      if (emit) {
        // first conv (unet.py:484): im2col of the (scaled) fp32 input into scrB, then a K=64 1x1 implicit GEMM
        kdip_unet* uu = u; int n = N, hh = H;
        F.push_back([=](cudaStream_t s) { return launch_im2col3x3(uu->io_x, uu->io_xscale, n, 3, hh, hh, scrB, s); });
        kdip_conv_desc d0;
        memset(&d0, 0, sizeof(d0));
        d0.N = N; d0.H = H; d0.W = H; d0.Cout_pad = b.cout; d0.Cout = b.cout; d0.nseg = 1;
        d0.seg[0].act = scrB; d0.seg[0].C = 64; d0.seg[0].wgt = u->w_in_i2c; d0.seg[0].taps = 1;
        d0.bias = u->b_in; d0.out = h.p; d0.out_mode = 0; d0.out_scale = 1.f;
        if (fused_stats) d0.chan_stats = h.stats;
        add_conv_op(F, conv(d0));
      }
      stats_op(F, h);
      rec.in0 = h;
    } else if (b.kind == 1) {
      const ResWeights& rw = u->resw[i];
      SavedRes sr;
      sr.two = (b.stage == 2 && b.first_of_block);
      sr.src0 = h;
      if (sr.two) { sr.src1 = hs.back(); rec.pop_id = (int)hs.size() - 1; hs.pop_back(); } else { sr.src1.p = nullptr; sr.src1.stats = nullptr; sr.src1.C = 0; sr.src1.H = sr.src1.W = 0; }
      const int C0 = sr.src0.C, C1 = sr.src1.C;
      if (C0 + C1 != b.cin) { set_error("unet plan: channel mismatch at %s (%d+%d != %d)", b.prefix.c_str(), C0, C1, b.cin); return KDIP_ESHAPE; }
      sr.Hin = sr.Win = H;
      const int rs = b.updown == 1 ? RS_AVGPOOL2 : (b.updown == 2 ? RS_NEAREST_UP2 : RS_NONE);
      if (b.updown == 1) H /= 2; else if (b.updown == 2) H *= 2;
      sr.Ho = sr.Wo = H;
      sr.ab1 = (float*)B.alloc((size_t)N * b.cin * 2 * 4);
      sr.mr1 = (float*)B.alloc((size_t)N * 32 * 2 * 4);
      sr.ab2 = (float*)B.alloc((size_t)N * b.cout * 2 * 4);
      sr.mr2 = (float*)B.alloc((size_t)N * 32 * 2 * 4);
      // Downsample blocks: the pooled identity skip avg_pool2d(x) is a by-product of the GroupNorm-apply pass
      bf16* xpool = (b.updown == 1 && b.cin == b.cout) ? (bf16*)B.alloc((size_t)N * H * H * b.cout * 2) : nullptr;
      sr.h1 = B.new_act(H, H, b.cout);
      sr.h1.stats = take_stats(b.cout);
      sr.out = B.new_act(H, H, b.cout);
      sr.out.stats = take_stats(b.cout);
      if (emit) {
        int n = N, Hin = sr.Hin, P_in = sr.Hin * sr.Win, Ho = H, Po = H * H, cin = b.cin, cout = b.cout, R = u->R;
        Act s0 = sr.src0, s1 = sr.src1, h1 = sr.h1;
        float *ab1 = sr.ab1, *mr1 = sr.mr1, *ab2 = sr.ab2, *mr2 = sr.mr2;
        const ResWeights w = rw;
        F.push_back([=](cudaStream_t s) {
          return launch_gn_finalize(s0.stats, s0.C, s1.stats, s1.C, n, P_in, w.g1, w.b1, nullptr, 0, 0, ab1, mr1, s);
        });
        kdip_conv_desc d1;
        memset(&d1, 0, sizeof(d1));
        d1.N = N; d1.H = Ho; d1.W = Ho; d1.Cout_pad = cout; d1.Cout = cout; d1.nseg = 1;
        d1.seg[0].act = scrA; d1.seg[0].C = cin; d1.seg[0].wgt = w.w1; d1.seg[0].taps = 9;
        d1.bias = w.bias1; d1.out = h1.p; d1.out_mode = 0; d1.out_scale = 1.f;
        if (fused_stats) d1.chan_stats = h1.stats;
        // GroupNorm + SiLU of in_layers (unet.py:237-243) on the conv's operand path where the row-tile pipeline runs: the
        // normalised tensor is never written; a concatenated input is two 3x3 K-segments over the raw sources
        bool fuse1 = false;
        if (rs == RS_NONE && (s1.C == 0 || (w.w1s0 != nullptr && w.w1s1 != nullptr))) {
          kdip_conv_desc df = d1;
          df.seg[0].act = s0.p; df.seg[0].C = s0.C; df.seg[0].wgt = s1.C > 0 ? w.w1s0 : w.w1;
          df.in_ab[0] = ab1;
          if (s1.C > 0) {
            df.seg[1].act = s1.p; df.seg[1].C = s1.C; df.seg[1].wgt = w.w1s1; df.seg[1].taps = 9;
            df.in_ab[1] = ab1 + (size_t)s0.C * 2;
            df.nseg = 2;
          }
          df.in_ab_C = cin; df.in_silu = 1;
          if (conv_can_fuse_gn_apply(&df)) { d1 = df; fuse1 = true; }
        }
        if (!fuse1) F.push_back([=](cudaStream_t s) { return launch_gn_apply(s0.p, s0.C, s1.p, s1.C, n, Hin, Hin, ab1, 1, rs, scrA, s, xpool); });
        add_conv_op(F, conv(d1));
        stats_op(F, h1);
        F.push_back([=](cudaStream_t s) {
          return launch_gn_finalize(h1.stats, cout, nullptr, 0, n, Po, w.g2, w.b2, film, R, w.film_off, ab2, mr2, s);
        });
        kdip_conv_desc d2;
        memset(&d2, 0, sizeof(d2));
        d2.N = N; d2.H = Ho; d2.W = Ho; d2.Cout_pad = cout; d2.Cout = cout;
        d2.seg[0].act = scrB; d2.seg[0].C = cout; d2.seg[0].wgt = w.w2; d2.seg[0].taps = 9;
        d2.nseg = 1;
        if (cin != cout) {
          d2.seg[1].act = s0.p; d2.seg[1].C = s0.C; d2.seg[1].wgt = w.ws0; d2.seg[1].taps = 1;
          d2.nseg = 2;
          if (s1.C > 0) { d2.seg[2].act = s1.p; d2.seg[2].C = s1.C; d2.seg[2].wgt = w.ws1; d2.seg[2].taps = 1; d2.nseg = 3; }
        } else {
          d2.residual = xpool ? (const void*)xpool : (const void*)s0.p;
          d2.res_mode = b.updown == 1 ? (xpool ? 1 : 2) : (b.updown == 2 ? 3 : 1);
        }
        d2.bias = w.bias2; d2.out = sr.out.p; d2.out_mode = 0; d2.out_scale = 1.f;
        if (fused_stats) d2.chan_stats = sr.out.stats;
        // GroupNorm + FiLM + SiLU of out_layers (unet.py:244-254) on the operand path (the skip segments stay raw)
        bool fuse2 = false;
        {
          kdip_conv_desc df = d2;
          df.seg[0].act = h1.p;
          df.in_ab[0] = ab2; df.in_ab_C = cout; df.in_silu = 1;
          if (conv_can_fuse_gn_apply(&df)) { d2 = df; fuse2 = true; }
        }
        if (!fuse2) F.push_back([=](cudaStream_t s) { return launch_gn_apply(h1.p, cout, nullptr, 0, n, Ho, Ho, ab2, 1, RS_NONE, scrB, s); });
        add_conv_op(F, conv(d2));
      }
      stats_op(F, sr.out);
      h = sr.out;
      rec.r = sr;
    } else {
      const AttnWeights& aw = u->attw[i];
      SavedAttn sa;
      sa.x = h;
      const int c = b.cin, T = H * H, heads = c / 64;
      sa.ab = (float*)B.alloc((size_t)N * c * 2 * 4);
      sa.mr = (float*)B.alloc((size_t)N * 32 * 2 * 4);
      sa.lse = (float*)B.alloc((size_t)N * heads * T * 4);
      sa.qkv = B.new_act(H, H, 3 * c);
      sa.att = B.new_act(H, H, c);
      sa.out = B.new_act(H, H, c);
      sa.out.stats = take_stats(c);
      if (emit) {
        int n = N, hh = H;
        Act x = sa.x, qkv = sa.qkv, att = sa.att;
        float *ab = sa.ab, *mr = sa.mr, *lse = sa.lse;
        const AttnWeights w = aw;
        F.push_back([=](cudaStream_t s) { return launch_gn_finalize(x.stats, c, nullptr, 0, n, T, w.g, w.b, nullptr, 0, 0, ab, mr, s); });
        F.push_back([=](cudaStream_t s) { return launch_gn_apply(x.p, c, nullptr, 0, n, hh, hh, ab, 0, RS_NONE, scrA, s); });
        kdip_conv_desc dq;
        memset(&dq, 0, sizeof(dq));
        dq.N = N; dq.H = H; dq.W = H; dq.Cout_pad = 3 * c; dq.Cout = 3 * c; dq.nseg = 1;
        dq.seg[0].act = scrA; dq.seg[0].C = c; dq.seg[0].wgt = w.wqkv; dq.seg[0].taps = 1;
        dq.bias = w.bqkv; dq.out = qkv.p; dq.out_mode = 0; dq.out_scale = 1.f;
        add_conv_op(F, conv(dq));
        F.push_back([=](cudaStream_t s) { return launch_attention_fwd(qkv.p, n, T, heads, 64, att.p, lse, s); });
        kdip_conv_desc dp;
        memset(&dp, 0, sizeof(dp));
        dp.N = N; dp.H = H; dp.W = H; dp.Cout_pad = c; dp.Cout = c; dp.nseg = 1;
        dp.seg[0].act = att.p; dp.seg[0].C = c; dp.seg[0].wgt = w.wproj; dp.seg[0].taps = 1;
        dp.bias = w.bproj; dp.residual = x.p; dp.res_mode = 1; dp.out = sa.out.p; dp.out_mode = 0; dp.out_scale = 1.f;
        if (fused_stats) dp.chan_stats = sa.out.stats;
        add_conv_op(F, conv(dp));
      }
      stats_op(F, sa.out);
      h = sa.out;
      rec.t = sa;
    }
    if (b.stage == 0 && b.last_of_block) {
      hs.push_back(h);
      hs_grad.push_back((bf16*)B.alloc((size_t)N * h.H * h.W * h.C * 2));
      rec.pushes = true;
      rec.push_id = (int)hs.size() - 1;
    }
    recs.push_back(rec);
  }
  if (!hs.empty()) { set_error("unet plan: skip stack not empty at the end"); return KDIP_EINVAL; }

  // ---- head ----
  const int c0 = h.C;
  float* ab_head = (float*)B.alloc((size_t)N * c0 * 2 * 4);
  float* mr_head = (float*)B.alloc((size_t)N * 32 * 2 * 4);
  Act hlast = h;
  if (emit) {
    kdip_unet* uu = u;
    int n = N, hh = H, P = H * H;
    F.push_back([=](cudaStream_t s) { return launch_gn_finalize(hlast.stats, c0, nullptr, 0, n, P, uu->g_head, uu->be_head, nullptr, 0, 0, ab_head, mr_head, s); });
    F.push_back([=](cudaStream_t s) { return launch_gn_apply(hlast.p, c0, nullptr, 0, n, hh, hh, ab_head, 1, RS_NONE, scrA, s); });
  }
  // The head conv writes straight into the caller's fp32 NCHW buffer, whose address is only known per call: its plan is
  // rebuilt lazily when the output pointer changes (see kdip_unet_forward).
  // ================================= backward =================================
  // Walk the records in reverse.  `gin` holds d/d(output of the block being processed).
  if (emit) {
    kdip_unet* uu = u;
    int n = N;
    std::vector<Op>& K = u->bwd_ops;
    size_t red_cursor = 0;
    (void)red_cursor;
    bf16* gin = g0;       // gradient wrt current block output
    bf16* gfree = g1;     // where the block-input gradient will be written
    {
      // head: seed (fp32 NCHW, 6 ch) -> g_a (g2) via direct dgrad; GN(+SiLU) backward -> gin
      float* red = take_red(c0);
      int hh = H;
      // head input-gradient (unet.py:617): im2col of the fp32 seed into g3, then a K=64 1x1 implicit GEMM -> g2
      K.push_back([=](cudaStream_t s) {
        KDIP_CUDA(cudaMemsetAsync(uu->red_base, 0, uu->red_bytes, s));
        return launch_im2col3x3(uu->io_seed, nullptr, n, 6, hh, hh, g3, s);
      });
      {
        kdip_conv_desc dh;
        memset(&dh, 0, sizeof(dh));
        dh.N = N; dh.H = H; dh.W = H; dh.Cout_pad = c0; dh.Cout = c0; dh.nseg = 1;
        dh.seg[0].act = g3; dh.seg[0].C = 64; dh.seg[0].wgt = u->w_headd_i2c; dh.seg[0].taps = 1;
        dh.out = g2; dh.out_mode = 0; dh.out_scale = 1.f;
        dh.gn_x0 = hlast.p; dh.gn_C0 = c0; dh.gn_silu = 1; dh.gn_ab = ab_head; dh.gn_red = red;
        const bool fused = conv_can_fuse_gn_reduce(&dh);
        if (!fused) dh.gn_red = nullptr;
        add_conv_op(K, conv(dh));
        if (!fused) K.push_back([=](cudaStream_t s) { return launch_gn_bwd_reduce(hlast.p, c0, nullptr, 0, n, hh, hh, ab_head, 1, RS_NONE, g2, red, s); });
      }
      K.push_back([=](cudaStream_t s) { return launch_gn_bwd_finalize(red, ab_head, mr_head, nullptr, n, c0, hh * hh, nullptr, 0, 0, kbuf, s); });
      K.push_back([=](cudaStream_t s) { return launch_gn_bwd_apply(hlast.p, c0, nullptr, 0, n, hh, hh, ab_head, kbuf, 1, RS_NONE, g2, nullptr, 0, gin, nullptr, s); });
    }
    auto sr_two = [](const Rec& r) { return r.kind == 1 && r.r.two; };
    bool skip_fused = false;   // the gradient `gin` already contains the skip-stack gradient of this block's output
    for (int ri = (int)recs.size() - 1; ri >= 0; --ri) {
      const Rec& rec = recs[ri];
      const BlockDesc& b = u->plan[rec.idx];
      // The block input (= previous block's output) may have been pushed on the skip stack: its skip gradient is added by this
      // block's GroupNorm-backward apply (one more addend) instead of a separate add pass over the tensor.
      const bf16* in_skip = (ri > 0 && recs[ri - 1].pushes && !sr_two(recs[ri]) && getenv("KDIP_UNFUSED_SKIPADD") == nullptr)
                                ? hs_grad[recs[ri - 1].push_id] : nullptr;
      if (rec.pushes && !skip_fused) {
        // this block's output was also consumed through the skip stack: add that gradient
        bf16* gs = hs_grad[rec.push_id];
        Act o = (rec.kind == 1) ? rec.r.out : (rec.kind == 2 ? rec.t.out : rec.in0);
        size_t cnt = (size_t)N * o.H * o.W * o.C;
        bf16* gi = gin;
        K.push_back([=](cudaStream_t s) { return launch_add_bf16(gi, gs, cnt, s); });
      }
      if (rec.kind == 1) {
        const SavedRes& sr = rec.r;
        const ResWeights w = u->resw[rec.idx];
        const int cin = b.cin, cout = b.cout, Ho = sr.Ho, Hin = sr.Hin;
        const int rs = b.updown == 1 ? RS_AVGPOOL2 : (b.updown == 2 ? RS_NEAREST_UP2 : RS_NONE);
        bf16* gi = gin;
        // dgrad conv2: g_out -> g_a2 (g2)
        kdip_conv_desc d;
        memset(&d, 0, sizeof(d));
        d.N = N; d.H = Ho; d.W = Ho; d.Cout_pad = cout; d.Cout = cout; d.nseg = 1;
        d.seg[0].act = gi; d.seg[0].C = cout; d.seg[0].wgt = w.w2d; d.seg[0].taps = 9;
        d.out = g2; d.out_mode = 0; d.out_scale = 1.f;
        // GN2 backward: (h1, g_a2) -> g_h1 (g3); its reduction pass rides in the dgrad conv's epilogue
        float* red2 = take_red(cout);
        Act h1 = sr.h1;
        float *ab2 = sr.ab2, *mr2 = sr.mr2;
        d.gn_x0 = h1.p; d.gn_C0 = cout; d.gn_silu = 1; d.gn_ab = ab2; d.gn_red = red2;
        const bool fused2 = conv_can_fuse_gn_reduce(&d);
        if (!fused2) d.gn_red = nullptr;
        add_conv_op(K, conv(d));
        if (!fused2) K.push_back([=](cudaStream_t s) { return launch_gn_bwd_reduce(h1.p, cout, nullptr, 0, n, Ho, Ho, ab2, 1, RS_NONE, g2, red2, s); });
        K.push_back([=](cudaStream_t s) { return launch_gn_bwd_finalize(red2, ab2, mr2, nullptr, n, cout, Ho * Ho, nullptr, 0, 0, kbuf, s); });
        K.push_back([=](cudaStream_t s) { return launch_gn_bwd_apply(h1.p, cout, nullptr, 0, n, Ho, Ho, ab2, kbuf, 1, RS_NONE, g2, nullptr, 0, g3, nullptr, s); });
        // dgrad conv1: g_h1 -> g_a1 (g2), Cin channels at the output resolution
        memset(&d, 0, sizeof(d));
        d.N = N; d.H = Ho; d.W = Ho; d.Cout_pad = cin; d.Cout = cin; d.nseg = 1;
        d.seg[0].act = g3; d.seg[0].C = cout; d.seg[0].wgt = w.w1d; d.seg[0].taps = 9;
        d.out = g2; d.out_mode = 0; d.out_scale = 1.f;
        float* red1 = take_red(cin);
        bool fused1 = false;
        if (rs == RS_NONE) {     // same-resolution GroupNorm input: fuse the reduction of GN1's backward
          d.gn_x0 = sr.src0.p; d.gn_C0 = sr.src0.C; d.gn_x1 = sr.src1.p; d.gn_silu = 1; d.gn_ab = sr.ab1; d.gn_red = red1;
          fused1 = conv_can_fuse_gn_reduce(&d);
          if (!fused1) d.gn_red = nullptr;
        }
        add_conv_op(K, conv(d));
        // skip path
        const bf16* extra = gi;
        int extra_mode = 2;
        if (cin != cout) {
          memset(&d, 0, sizeof(d));
          d.N = N; d.H = Ho; d.W = Ho; d.Cout_pad = cin; d.Cout = cin; d.nseg = 1;
          d.seg[0].act = gi; d.seg[0].C = cout; d.seg[0].wgt = w.wsd; d.seg[0].taps = 1;
          d.out = g4; d.out_mode = 0; d.out_scale = 1.f;
          add_conv_op(K, conv(d));
          extra = g4;
          extra_mode = 1;
        }
        // GN1 backward (+resample^T) + skip gradient -> gradients of the (one or two) sources
        Act s0 = sr.src0, s1 = sr.src1;
        float *ab1 = sr.ab1, *mr1 = sr.mr1;
        bf16* dst0 = gfree;
        bf16* dst1 = sr.two ? hs_grad[rec.pop_id] : nullptr;
        if (!fused1) K.push_back([=](cudaStream_t s) { return launch_gn_bwd_reduce(s0.p, s0.C, s1.p, s1.C, n, Hin, Hin, ab1, 1, rs, g2, red1, s); });
        K.push_back([=](cudaStream_t s) { return launch_gn_bwd_finalize(red1, ab1, mr1, nullptr, n, cin, Hin * Hin, nullptr, 0, 0, kbuf, s); });
        K.push_back([=](cudaStream_t s) {
          return launch_gn_bwd_apply(s0.p, s0.C, s1.p, s1.C, n, Hin, Hin, ab1, kbuf, 1, rs, g2, extra, extra_mode, dst0, dst1, s, in_skip);
        });
        skip_fused = in_skip != nullptr;
        std::swap(gin, gfree);
      } else if (rec.kind == 2) {
        const SavedAttn& sa = rec.t;
        const AttnWeights w = u->attw[rec.idx];
        const int c = b.cin, hh = sa.x.H, T = hh * hh, heads = c / 64;
        bf16* gi = gin;
        kdip_conv_desc d;
        memset(&d, 0, sizeof(d));
        d.N = N; d.H = hh; d.W = hh; d.Cout_pad = c; d.Cout = c; d.nseg = 1;
        d.seg[0].act = gi; d.seg[0].C = c; d.seg[0].wgt = w.wprojd; d.seg[0].taps = 1;
        d.out = g2; d.out_mode = 0; d.out_scale = 1.f;
        add_conv_op(K, conv(d));                                     // g_att (g2)
        Act qkv = sa.qkv, att = sa.att, x = sa.x;
        float *ab = sa.ab, *mr = sa.mr, *lse = sa.lse;
        K.push_back([=](cudaStream_t s) { return launch_attention_bwd(qkv.p, att.p, g2, lse, n, T, heads, 64, g3, s); });   // g_qkv (g3)
        memset(&d, 0, sizeof(d));
        d.N = N; d.H = hh; d.W = hh; d.Cout_pad = c; d.Cout = c; d.nseg = 1;
        d.seg[0].act = g3; d.seg[0].C = 3 * c; d.seg[0].wgt = w.wqkvd; d.seg[0].taps = 1;
        d.out = g2; d.out_mode = 0; d.out_scale = 1.f;
        float* red = take_red(c);
        d.gn_x0 = x.p; d.gn_C0 = c; d.gn_silu = 0; d.gn_ab = ab; d.gn_red = red;
        const bool fuseda = conv_can_fuse_gn_reduce(&d);
        if (!fuseda) d.gn_red = nullptr;
        add_conv_op(K, conv(d));                                     // g_a (g2)
        bf16* dst0 = gfree;
        if (!fuseda) K.push_back([=](cudaStream_t s) { return launch_gn_bwd_reduce(x.p, c, nullptr, 0, n, hh, hh, ab, 0, RS_NONE, g2, red, s); });
        K.push_back([=](cudaStream_t s) { return launch_gn_bwd_finalize(red, ab, mr, nullptr, n, c, T, nullptr, 0, 0, kbuf, s); });
        K.push_back([=](cudaStream_t s) { return launch_gn_bwd_apply(x.p, c, nullptr, 0, n, hh, hh, ab, kbuf, 0, RS_NONE, g2, gi, 1, dst0, nullptr, s, in_skip); });
        skip_fused = in_skip != nullptr;
        std::swap(gin, gfree);
      } else {
        // conv_in dgrad: g_h0 -> fp32 NCHW grad wrt the (scaled) UNet input; plan rebuilt lazily on pointer change
        u->gin_final = gin;
      }
    }
    if (red_used > red_floats) { set_error("unet plan: GN-backward reduction region overflow"); return KDIP_EINVAL; }
  }
  if (stats_used > stats_floats) { set_error("unet plan: statistics region overflow"); return KDIP_EINVAL; }
  if (rc != KDIP_OK) return rc;
  *need_bytes = B.cur;
  if (emit) {
    if (B.cur > ws_bytes) { set_error("unet: workspace too small: need %zu bytes, got %zu", B.cur, ws_bytes); return KDIP_ENOMEM; }
    u->planned_N = N; u->planned_ws = ws; u->planned_bytes = ws_bytes;
    u->scrA_ptr = scrA;
    u->scrB_ptr = scrB;
    u->g3_ptr = g3;
    u->hlast_ptr = hlast.p;
  }
  return KDIP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// state_dict schema (names and shapes of UNetModel's parameters, unet.py:463-618) derived from the same block plan
// ---------------------------------------------------------------------------------------------------------------------
namespace kdip {
struct SchemaEntry { std::string name; int64_t shape[4]; int ndim; };
static void build_schema(const kdip_unet_arch& a, std::vector<SchemaEntry>& out) {
  std::vector<BlockDesc> plan;
  build_block_plan(a, plan);
  auto add = [&](const std::string& n, std::initializer_list<int64_t> shp) {
    SchemaEntry e;
    e.name = n; e.ndim = (int)shp.size();
    int i = 0;
    for (int64_t v : shp) e.shape[i++] = v;
    for (; i < 4; ++i) e.shape[i] = 1;
    out.push_back(e);
  };
  const int64_t mc = a.model_channels, ted = 4 * mc;
  add("time_embed.0.weight", {ted, mc}); add("time_embed.0.bias", {ted});
  add("time_embed.2.weight", {ted, ted}); add("time_embed.2.bias", {ted});
  for (auto& b : plan) {
    const std::string& p = b.prefix;
    const int64_t ci = b.cin, co = b.cout;
    if (b.kind == 0) {
      add(p + ".weight", {co, ci, 3, 3}); add(p + ".bias", {co});
    } else if (b.kind == 1) {
      add(p + ".in_layers.0.weight", {ci}); add(p + ".in_layers.0.bias", {ci});
      add(p + ".in_layers.2.weight", {co, ci, 3, 3}); add(p + ".in_layers.2.bias", {co});
      add(p + ".emb_layers.1.weight", {2 * co, ted}); add(p + ".emb_layers.1.bias", {2 * co});
      add(p + ".out_layers.0.weight", {co}); add(p + ".out_layers.0.bias", {co});
      add(p + ".out_layers.3.weight", {co, co, 3, 3}); add(p + ".out_layers.3.bias", {co});
      if (ci != co) { add(p + ".skip_connection.weight", {co, ci, 1, 1}); add(p + ".skip_connection.bias", {co}); }
    } else {
      add(p + ".norm.weight", {ci}); add(p + ".norm.bias", {ci});
      add(p + ".qkv.weight", {3 * ci, ci, 1}); add(p + ".qkv.bias", {3 * ci});
      add(p + ".proj_out.weight", {ci, ci, 1}); add(p + ".proj_out.bias", {ci});
    }
  }
  const int64_t c0 = (int64_t)(a.channel_mult[0] * mc);
  add("out.0.weight", {c0}); add("out.0.bias", {c0});
  add("out.2.weight", {a.out_channels, c0, 3, 3}); add("out.2.bias", {a.out_channels});
}
}  // namespace kdip

extern "C" int kdip_unet_schema_count(const kdip_unet_arch* arch, int* n) {
  KDIP_REQUIRE(arch && n, KDIP_EINVAL, "unet_schema_count: null argument");
  std::vector<SchemaEntry> sc;
  build_schema(*arch, sc);
  *n = (int)sc.size();
  return KDIP_OK;
}
extern "C" int kdip_unet_schema_entry(const kdip_unet_arch* arch, int index, char* name_out, int name_cap, int64_t* shape_out,
                                      int* ndim) {
  KDIP_REQUIRE(arch && name_out && shape_out && ndim && name_cap > 0, KDIP_EINVAL, "unet_schema_entry: null argument");
  std::vector<SchemaEntry> sc;
  build_schema(*arch, sc);
  KDIP_REQUIRE(index >= 0 && index < (int)sc.size(), KDIP_EINVAL, "unet_schema_entry: index %d out of range", index);
  snprintf(name_out, (size_t)name_cap, "%s", sc[index].name.c_str());
  for (int i = 0; i < 4; ++i) shape_out[i] = sc[index].shape[i];
  *ndim = sc[index].ndim;
  return KDIP_OK;
}

// pre-head feature of the last forward as fp32 NCHW [N, C0, S, S] (UNetModel.forward(return_feature=True), unet.py:665-666)
extern "C" int kdip_unet_feature(kdip_unet* u, int N, float* feat, kdip_stream_t s) {
  KDIP_REQUIRE(u && feat, KDIP_EINVAL, "unet_feature: null argument");
  if (u->f32) return fp32_feature(u->f32, N, feat, (cudaStream_t)s);
  KDIP_REQUIRE(u->planned_N == N && u->hlast_ptr, KDIP_EINVAL, "unet_feature: must follow kdip_unet_forward with the same N");
  const int S = u->arch.image_size, c0 = (int)(u->arch.channel_mult[0] * u->arch.model_channels);
  return kdip_nhwc_bf16_to_nchw_f32(u->hlast_ptr, N, c0, S, S, feat, s);
}

extern "C" int kdip_unet_workspace_bytes(kdip_unet* u, int N, size_t* bytes) {
  KDIP_REQUIRE(u && bytes && N > 0, KDIP_EINVAL, "unet_workspace_bytes: bad argument");
  if (u->f32) return fp32_workspace_bytes(u->f32, N, bytes);
  return build_launch_plan(u, N, nullptr, 0, bytes);
}

static int ensure_plan(kdip_unet* u, int N, void* ws, size_t ws_bytes) {
  KDIP_REQUIRE(ws != nullptr && ((uintptr_t)ws % 256) == 0, KDIP_EALIGN, "unet: workspace must be 256-byte aligned");
  if (u->planned_N == N && u->planned_ws == ws && u->planned_bytes == ws_bytes) return KDIP_OK;
  u->planned_N = 0;
  u->head_plan = nullptr; u->cov_plan = nullptr; u->ind_plan = nullptr;
  size_t need = 0;
  return build_launch_plan(u, N, ws, ws_bytes, &need);
}

extern "C" int kdip_unet_prepare(kdip_unet* u, int N, void* workspace, size_t ws_bytes) {
  KDIP_REQUIRE(u && N > 0, KDIP_EINVAL, "unet_prepare: bad argument");
  if (u->f32) return fp32_prepare(u->f32, N, workspace, ws_bytes);
  return ensure_plan(u, N, workspace, ws_bytes);
}

// drop a cached per-pointer plan that is about to be replaced
static void retire_plan(kdip_unet* u, ConvPlan* old) {
  if (!old) return;
  for (size_t i = 0; i < u->conv_plans.size(); ++i)
    if (u->conv_plans[i] == old) {
      u->conv_plans.erase(u->conv_plans.begin() + i);
      break;
    }
  conv_plan_free(old);
}

// head conv(s) into the caller's buffers (plans cached per output pointer)
static int forward_tail(kdip_unet* u, int N, float* out, float* cov_out, cudaStream_t s, double* flops) {
  int rc;
  const int S = u->arch.image_size, c0 = (int)(u->arch.channel_mult[0] * u->arch.model_channels);
  if (flops) *flops = 2.0 * N * S * S * 6.0 * c0 * 9;
  if (u->head_plan == nullptr) {   // output goes to scratch: the plan no longer depends on the caller's pointer
    kdip_conv_desc d;
    memset(&d, 0, sizeof(d));
    // head conv3x3 C0 -> 6 (unet.py:617) as a tap-folded 1x1 GEMM (N = 54 -> 64) into scratch + 9-neighbour gather
    d.N = N; d.H = S; d.W = S; d.Cout_pad = 64; d.Cout = 64; d.nseg = 1;
    d.seg[0].act = u->scrA_ptr; d.seg[0].C = c0; d.seg[0].wgt = u->w_head_fold; d.seg[0].taps = 1;
    d.out = u->scrB_ptr; d.out_mode = 2; d.out_scale = 1.f;
    ConvPlan* p = make_conv(u, d, &rc);
    if (rc != KDIP_OK) return rc;
    retire_plan(u, u->head_plan);
    u->head_plan = p; u->head_out = out;
  }
  rc = conv_plan_launch(u->head_plan, s);
  if (rc != KDIP_OK) return rc;
  rc = launch_tap_gather((const float*)u->scrB_ptr, 64, u->b_head, N, 6, S, S, out, s);
  if (rc != KDIP_OK) return rc;
  if (cov_out) {
    if (u->cov_plan == nullptr || u->cov_out_ptr != cov_out) {
      kdip_conv_desc d;
      memset(&d, 0, sizeof(d));
      d.N = N; d.H = S; d.W = S; d.Cout_pad = 16; d.Cout = 6; d.nseg = 1;
      d.seg[0].act = u->hlast_ptr; d.seg[0].C = c0; d.seg[0].wgt = u->w_cov; d.seg[0].taps = 1;
      d.bias = u->b_cov; d.out = cov_out; d.out_mode = 1; d.out_scale = 1.f;
      ConvPlan* p = make_conv(u, d, &rc);
      if (rc != KDIP_OK) return rc;
      retire_plan(u, u->cov_plan);
      u->cov_plan = p; u->cov_out_ptr = cov_out;
    }
    rc = conv_plan_launch(u->cov_plan, s);
    if (rc != KDIP_OK) return rc;
  }
  return KDIP_OK;
}

extern "C" int kdip_unet_forward(kdip_unet* u, const float* x, const float* x_scale, const float* t, int N, float* out,
                                 float* cov_out, void* workspace, size_t ws_bytes, kdip_stream_t stream) {
  KDIP_REQUIRE(u && x && t && out && N > 0, KDIP_EINVAL, "unet_forward: bad argument");
  KDIP_REQUIRE(cov_out == nullptr || u->has_cov, KDIP_EINVAL, "unet_forward: cov_out requested but no out_cov weights were given");
  if (u->f32) return fp32_forward(u->f32, x, x_scale, t, N, out, cov_out, workspace, ws_bytes, (cudaStream_t)stream);
  int rc = ensure_plan(u, N, workspace, ws_bytes);
  if (rc != KDIP_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  u->io_x = x; u->io_xscale = x_scale; u->io_t = t; u->io_out = out; u->io_cov = cov_out;
  for (auto& op : u->fwd_ops) {
    rc = op(s);
    if (rc != KDIP_OK) return rc;
  }
  return forward_tail(u, N, out, cov_out, s, nullptr);
}

// conv_in input-gradient into the caller's buffer
static int vjp_tail(kdip_unet* u, int N, float* grad_x, cudaStream_t s, double* flops) {
  int rc;
  const int S = u->arch.image_size, c0 = (int)(u->arch.channel_mult[0] * u->arch.model_channels);
  if (flops) *flops = 2.0 * N * S * S * 3.0 * c0 * 9;
  if (u->ind_plan == nullptr) {
    kdip_conv_desc d;
    memset(&d, 0, sizeof(d));
    // first layer's input-gradient C0 -> 3 as a tap-folded 1x1 GEMM (N = 27 -> 32) into scratch + 9-neighbour gather
    d.N = N; d.H = S; d.W = S; d.Cout_pad = 32; d.Cout = 32; d.nseg = 1;
    d.seg[0].act = u->gin_final; d.seg[0].C = c0; d.seg[0].wgt = u->w_ind_fold; d.seg[0].taps = 1;
    d.out = u->g3_ptr; d.out_mode = 2; d.out_scale = 1.f;
    ConvPlan* p = make_conv(u, d, &rc);
    if (rc != KDIP_OK) return rc;
    retire_plan(u, u->ind_plan);
    u->ind_plan = p; u->ind_out = grad_x;
  }
  rc = conv_plan_launch(u->ind_plan, s);
  if (rc != KDIP_OK) return rc;
  return launch_tap_gather((const float*)u->g3_ptr, 32, nullptr, N, 3, S, S, grad_x, s);
}

extern "C" int kdip_unet_vjp(kdip_unet* u, const float* seed, int N, float* grad_x, void* workspace, size_t ws_bytes,
                             kdip_stream_t stream) {
  KDIP_REQUIRE(u && seed && grad_x && N > 0, KDIP_EINVAL, "unet_vjp: bad argument");
  if (u->f32) return fp32_vjp(u->f32, seed, N, grad_x, workspace, ws_bytes, (cudaStream_t)stream);
  KDIP_REQUIRE(u->planned_N == N && u->planned_ws == workspace && u->planned_bytes == ws_bytes, KDIP_EINVAL,
               "unet_vjp: must follow kdip_unet_forward with the same N and workspace (saved activations live there)");
  cudaStream_t s = (cudaStream_t)stream;
  u->io_seed = seed; u->io_grad = grad_x;
  int rc;
  for (auto& op : u->bwd_ops) {
    rc = op(s);
    if (rc != KDIP_OK) return rc;
  }
  return vjp_tail(u, N, grad_x, s, nullptr);
}

extern "C" int kdip_unet_profile(kdip_unet* u, const float* x, const float* x_scale, const float* t, const float* seed, int N,
                                 float* out, float* grad_x, void* workspace, size_t ws_bytes, kdip_stream_t stream,
                                 kdip_unet_profile_t* prof) {
  KDIP_REQUIRE(u && x && t && out && seed && grad_x && prof && N > 0, KDIP_EINVAL, "unet_profile: bad argument");
  KDIP_REQUIRE(u->f32 == nullptr, KDIP_EINVAL, "unet_profile: only the bf16 tcgen05 engine is instrumented");
  int rc = ensure_plan(u, N, workspace, ws_bytes);
  if (rc != KDIP_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  u->io_x = x; u->io_xscale = x_scale; u->io_t = t; u->io_out = out; u->io_cov = nullptr;
  u->io_seed = seed; u->io_grad = grad_x;
  // One event BETWEEN consecutive launches (the end of step i is the start of step i + 1), all created up front: a pair per step put
  // two event records into every gap and inflated the per-launch times by ~20 us each (26.4 ms of "conv" against 22.8 ms in ncu).
  struct Rec { bool conv; double flops; int e0, e1; };
  std::vector<Rec> recs;
  const size_t n_steps = u->fwd_ops.size() + u->bwd_ops.size() + 2;
  std::vector<cudaEvent_t> ev(n_steps + 3);
  for (auto& e : ev) KDIP_CUDA(cudaEventCreate(&e));
  int cur = 0;                                     // index of the most recently recorded event
  KDIP_CUDA(cudaEventRecord(ev[0], s));
  auto mark = [&]() -> int { KDIP_CUDA(cudaEventRecord(ev[cur + 1], s)); ++cur; return KDIP_OK; };
  auto timed = [&](bool is_conv, double flops, const std::function<int()>& fn) -> int {
    const int e0 = cur;
    int q = fn();
    if (q != KDIP_OK) return q;
    q = mark();
    recs.push_back(Rec{is_conv, flops, e0, cur});
    return q;
  };
  double fl = 0.0;
  for (auto& op : u->fwd_ops) {
    rc = timed(op.is_conv, op.flops, [&]() { return op(s); });
    if (rc != KDIP_OK) return rc;
  }
  forward_tail(u, N, out, nullptr, s, &fl);   // compute flops first (plan cached on the second call)
  rc = mark();                                // ... untimed: the next step starts after it
  if (rc != KDIP_OK) return rc;
  rc = timed(true, fl, [&]() { return forward_tail(u, N, out, nullptr, s, nullptr); });
  if (rc != KDIP_OK) return rc;
  for (auto& op : u->bwd_ops) {
    rc = timed(op.is_conv, op.flops, [&]() { return op(s); });
    if (rc != KDIP_OK) return rc;
  }
  vjp_tail(u, N, grad_x, s, &fl);
  rc = mark();
  if (rc != KDIP_OK) return rc;
  rc = timed(true, fl, [&]() { return vjp_tail(u, N, grad_x, s, nullptr); });
  if (rc != KDIP_OK) return rc;
  KDIP_CUDA(cudaStreamSynchronize(s));
  memset(prof, 0, sizeof(*prof));
  for (auto& r : recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[r.e0], ev[r.e1]);
    if (r.conv) { prof->conv_ms += ms; prof->conv_flops += r.flops; prof->conv_launches++; }
    else { prof->other_ms += ms; prof->other_steps++; }
    prof->total_ms += ms;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return KDIP_OK;
}
