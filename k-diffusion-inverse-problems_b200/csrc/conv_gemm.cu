// Implicit-GEMM convolution for the ADM UNet on the 5th-gen tensor cores (sm_100a).
//
// Replaces the cuDNN/cuBLAS calls under guided_diffusion/unet.py:185,211 (ResBlock conv3x3), :222 (1x1 skip),
// :287,295 (attention qkv / proj conv1d), :617 (output head) and their input-gradients (autograd sites
// condition/condition.py:136,146,155,172,269).
//
// GEMM view:  D[pixel, co] = sum_seg sum_tap sum_ci  act_seg[pixel + offset(tap), ci] * W_seg[tap][co][ci]
//   M = 128 output pixels per CTA tile (a TN x TH x TW box of the NHWC activation),
//   N = BN output channels (16..256), K = 64-channel blocks walked over (segment, tap, channel chunk).
// Data movement: one elected producer thread issues TMA tile loads — the A box is the *shifted* pixel box of the
//   tap (out-of-bounds rows/cols are zero-filled by TMA = the conv's zero padding), the B box is BN rows of the
//   tap's weight slab — into a ring of 128B-swizzled shared-memory stages.
// Math: one elected thread issues tcgen05.mma (M=128, N=BN, K=16, bf16 x bf16 -> fp32) with the accumulator in
//   TMEM, double-buffered so the epilogue of tile i overlaps the main loop of tile i+1 (persistent CTAs).
// Epilogue, bf16 NHWC outputs with BN >= 64 (every ResBlock / attention conv): 8 warps, two per TMEM lane quadrant, each
//   owning a 64-channel slab.  tcgen05.ld (2 x 32 columns in flight) -> +bias (shared memory) -> + identity residual (its
//   tile is TMA-loaded by a dedicated producer thread, unet.py:257,306) -> bf16 -> 128B-swizzled staging tile in shared
//   memory -> TMA store (cp.async.bulk.tensor, full 128 B lines; no scattered per-lane sectors).
//   Optional fused GroupNorm statistics: per-(image, channel) sum / sum of squares of the stored tile, read back from the
//   staging tile (conflict-free) and accumulated with 256 fp32 REDs per 128-channel pass.
// Legacy epilogue (4 warps, per-lane global loads/stores) for the rare shapes: fp32 NCHW outputs with <= 32 channels (head,
//   input gradient), 2x2 avg-pool / nearest-up skip paths (unet.py:190-197).
#include <stdlib.h>

#define KDIP_MBAR_WAIT_QUIET
#include "kdip_common.cuh"

namespace kdip {

// A protocol bug must be diagnosable from the host: a wait that times out leaves (source line, block, thread, parity, barrier
// address) in a host-mapped record before it traps (the context is dead afterwards, mapped host memory is not).  Read with
// kdip_conv_trap_read.  No printf: the control warps run on a 48-register budget and cannot afford its call frame.
static __device__ unsigned int* g_trap_rec = nullptr;     // [0] = number of records, then 4 words per record (up to 15)
static unsigned int* g_trap_host = nullptr;
__device__ __forceinline__ void mbar_wait_line(uint64_t* bar, uint32_t parity, uint32_t line) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      unsigned int* rec = g_trap_rec;
      if (rec != nullptr) {
        const unsigned int i = atomicAdd_system(rec, 1u);
        if (i < 15u) {
          volatile unsigned int* r = rec + 4 + 4 * i;
          r[0] = line; r[1] = blockIdx.x; r[2] = threadIdx.x | (parity << 16); r[3] = smem_u32(bar);
        }
        __threadfence_system();
      }
      __trap();
    }
  }
}
#define mbar_wait(bar, parity) mbar_wait_line(bar, parity, (uint32_t)__LINE__)

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;                      // bf16 elements = 128 bytes = one swizzle row
static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
static constexpr int kThreads = 384;                    // 4 control warps + 8 epilogue warps
static constexpr int kXfThreads = 640;                  // + 8 operand-transform warps (fused GroupNorm apply, halo pipeline only)
static constexpr int kEpiThreads = 256;
static constexpr int kSlabBytes = kBlockM * 128;        // one 64-channel bf16 slab of the output tile: 16 KiB
static constexpr int kStagingBytes = 2 * kSlabBytes;    // 128 channels at a time
static constexpr int kMaxStages = 8;
// "halo" pipeline (3x3 convs on images at least 128 pixels wide): a pixel tile is 128 consecutive pixels of ONE image row and a
// work item is two vertically adjacent tiles.  Per 64-channel chunk the producer loads the four input rows y-1..y+2 once, 130
// pixels wide (x0-1 .. x0+128, zero-filled outside the image), and all nine taps of both tiles read them through shared-memory
// descriptors that start dx+1 rows into the slot (the operand swizzle follows the absolute address): 66 KB of activations per chunk and work
// item instead of 2 x 9 x 16 KB.  Weights stream through their own ring, one (tap, chunk) slab per stage.
static constexpr int kHaloPix = 130;
static constexpr int kHaloRowBytes = kHaloPix * 128;    // 16640 B landed by TMA
static constexpr int kHaloSlot = 17 * 1024;             // slot stride (keeps every slot 1024-byte aligned)
static constexpr int kHaloSlots = 6;                    // default ring: four rows of the current chunk + two of the next
static constexpr int kMaxHaloSlots = 8;                 // ConvParams.halo_slots (KDIP_HALO_SLOTS) may deepen the ring

struct ConvParams {
  CUtensorMap mapA[3];
  CUtensorMap mapB[3];
  CUtensorMap mapOut;   // bf16 NHWC output, box {64, TW, TH, TN} (TMA-store epilogue)
  CUtensorMap mapRes;   // identity residual, same geometry; or the first GroupNorm source of the fused backward reduction
  CUtensorMap mapRes2;  // second (concatenated) GroupNorm source
  int tma_epilogue;     // 1: 8-warp TMA-store epilogue; 0: legacy 4-warp epilogue
  int pair;             // 1: CTA pairs (cta_group::2, M = 256 per pair, B split across the two CTAs)
  int mt;               // pixel tiles per work item (1 or 2): mt = 2 shares every weight stage between two M=128 accumulators
  int halo;             // 1: halo pipeline (see kHaloPix)
  int halo_slots;       // activation-row slots of the ring (kHaloSlots .. kMaxHaloSlots)
  int halo_bo;          // 1: descriptors carry the matrix base offset (probe only; wrong on B200, see conv_plan_build)
  int ws;               // 1: the two tiles of a work item share the weight operand through the tensor core's collector (tcgen05.mma.ws)
  int dbg;              // timing experiments only (KDIP_CONV_DBG): 1 = no operand loads / waits, 2 = epilogue releases TMEM without reading or storing,
                        // 4 = operand-transform warps only hand the rows over, 8 = they copy the rows through registers without the math,
                        // 16 = transform warpgroups wait only for their own rows (the protocol bug fixed in round 2, kept for the regression experiment)
  uint32_t res_slab_bytes;   // bytes TMA lands per 64-channel slab of the skip / GroupNorm-source tile
  int b_rows;           // weight rows each CTA loads per k-block: BN (single) or BN/2 (pair)
  int total_work;       // persistent-loop trip count: tiles (single) or pair tiles (pair)
  int seg_taps[3];
  int seg_chunks[3];
  int nseg;
  int N, H, W;
  int TW, TH, TN;
  int tiles_x, tiles_y, tiles_n, n_tiles, total_tiles;
  int BN, Cout_pad, Cout;
  int num_stages;
  uint32_t idesc;
  const float* bias;
  const __nv_bfloat16* residual;
  int res_mode;
  int res_tma;          // 1: the skip tile is TMA-loaded (res_mode 1; res_mode 3 with a half-size source box)
  void* out;
  int out_mode;
  float out_scale;
  float* chan_stats;
  // fused GroupNorm-backward reduction (kdip_conv_desc.gn_*): the x tile is TMA-loaded like a residual tile
  const float* gn_ab;
  float* gn_red;
  int gn_C0, gn_silu, gn_two;
  // fused GroupNorm apply on the operand path (kdip_conv_desc.in_ab): segment s is consumed as act(A x + B)
  const float* xf_ab[3];
  int xf_C, xf_silu;
};

struct TileCoord {
  int n0, y0, x0, nn0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int tile) {
  TileCoord t;
  int nt = tile % p.n_tiles;
  int mt = tile / p.n_tiles;
  t.nn0 = nt * p.BN;
  if (p.halo) {
    // pixel tiles 2q and 2q+1 are rows 2yp and 2yp+1 of the same 128-pixel column block
    const int j = mt & 1, q = mt >> 1;
    const int tx = q % p.tiles_x, r = q / p.tiles_x, hh = p.H >> 1;
    t.x0 = tx * 128;
    t.y0 = 2 * (r % hh) + j;
    t.n0 = r / hh;
    return t;
  }
  int tx = mt % p.tiles_x;
  int r = mt / p.tiles_x;
  int ty = r % p.tiles_y;
  int tn = r / p.tiles_y;
  t.x0 = tx * p.TW;
  t.y0 = ty * p.TH;
  t.n0 = tn * p.TN;
  return t;
}

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// work item -> tile index of this CTA's j-th pixel tile.  work = pm * n_tiles + nt.
//   single CTA: M-tiles pm*mt + j;   pair mode: the leader owns M-tiles 2pm*mt + j, the peer (2pm+1)*mt + j.
__device__ __forceinline__ int work_to_tile(const ConvParams& p, int work, int rank, int j) {
  const int nt = work % p.n_tiles, pm = work / p.n_tiles;
  const int m = (p.pair ? (2 * pm + rank) : pm) * p.mt + j;
  return m * p.n_tiles + nt;
}

// register budgets of the 640-thread variant (setmaxnreg per warpgroup): the CTA starts with 640 x 96 = 61440 registers;
// control warps 72, two epilogue warpgroups 136, two transform warpgroups 64: (72 + 2 x 136 + 2 x 64) x 128 = 60416
// Measured (tools/experiments/r2e_xf_regs.sh, 3x3 128->128 at 256^2, B=32): 48/72/144 550 us, 64/72/136 543 us, 72/64/136 538 us - the
// control warps (TMA producers, MMA issuer) are the ones that profit from registers; 40/56/160 and 48/64/152 were no better than 48/72/144.
#ifndef KDIP_XF_REG_CTL
#define KDIP_XF_REG_CTL 72
#define KDIP_XF_REG_XF 64
#define KDIP_XF_REG_EPI 136
#endif
#define KDIP_STR2(x) #x
#define KDIP_STR(x) KDIP_STR2(x)
static_assert(4 * KDIP_XF_REG_CTL + 8 * KDIP_XF_REG_XF + 8 * KDIP_XF_REG_EPI <= 20 * 96, "register budgets exceed the CTA's pool");
__device__ __forceinline__ void setmaxnreg_ctl() { asm volatile("setmaxnreg.dec.sync.aligned.u32 " KDIP_STR(KDIP_XF_REG_CTL) ";" ::: "memory"); }
__device__ __forceinline__ void setmaxnreg_xf() { asm volatile("setmaxnreg.dec.sync.aligned.u32 " KDIP_STR(KDIP_XF_REG_XF) ";" ::: "memory"); }
__device__ __forceinline__ void setmaxnreg_epi() { asm volatile("setmaxnreg.inc.sync.aligned.u32 " KDIP_STR(KDIP_XF_REG_EPI) ";" ::: "memory"); }

// kXf (halo pipeline only): eight more warps (two warpgroups taking alternate rows) apply the GroupNorm affine + SiLU of the conv's input (nn.py:17-19, unet.py:237-257) to the
// activation rows IN shared memory, between the TMA landing and the MMAs - the normalised tensor never exists in HBM.
template <bool kPair, bool kHalo, bool kXf>
__global__ void __launch_bounds__(kXf ? kXfThreads : kThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  static_assert(!kXf || kHalo, "the operand transform lives in the halo pipeline");
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms (TMA destination and UMMA descriptors)
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int mt = p.mt;
  const int a_bytes = mt * kABytes;
  const int stage_bytes = a_bytes + p.b_rows * 128;
  const int rank = kPair ? (int)cluster_ctarank() : 0;
  const int work0 = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_stride = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // halo mode: [kHaloSlots activation-row slots][num_stages weight stages of BN x 128 B]
  const int nslots = p.halo_slots;
  uint8_t* b_ring = smem + nslots * kHaloSlot;
  const int b_stage_bytes = p.b_rows * 128;
  uint8_t* staging = kHalo ? b_ring + p.num_stages * b_stage_bytes : smem + p.num_stages * stage_bytes;   // [2][128 rows][128 B], TMA-store source
  uint8_t* res_stage = staging + (p.tma_epilogue ? kStagingBytes : 0);        // residual tile, same layout
  uint8_t* after = res_stage + (p.res_tma ? kStagingBytes : 0);
  float* bias_s = reinterpret_cast<float*>(after);                            // [256]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(after + 1024);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kMaxStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_full_bar = tmem_empty_bar + 2;
  uint64_t* res_empty_bar = res_full_bar + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_empty_bar + 1);
  uint64_t* a_full_bar = res_empty_bar + 2;            // halo mode: activation-row slots
  uint64_t* a_empty_bar = a_full_bar + kMaxHaloSlots;
  uint64_t* a_ready_bar = a_empty_bar + kMaxHaloSlots;    // kXf: rows transformed (4 warps per CTA arrive; on the leader for pairs)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.mapA[s]);
      tma_prefetch_desc(&p.mapB[s]);
    }
    if (p.tma_epilogue) {
      tma_prefetch_desc(&p.mapOut);
      if (p.res_tma) tma_prefetch_desc(&p.mapRes);
      if (p.gn_two) tma_prefetch_desc(&p.mapRes2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(&full_bar[s], kPair ? 2 : 1);   // pair: one arrive.expect_tx per CTA, both on the leader's barrier
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kPair ? 2 * kEpiThreads : (p.tma_epilogue ? kEpiThreads : 128));
    }
    mbar_init(res_full_bar, 1);
    mbar_init(res_empty_bar, kEpiThreads);
    if (kHalo) {
      for (int s = 0; s < nslots; ++s) {
        // pair: one arrive.expect_tx per CTA, both on the leader's barrier; kXf: every CTA's rows land on its own barrier
        mbar_init(&a_full_bar[s], (kPair && !kXf) ? 2 : 1);
        mbar_init(&a_empty_bar[s], 1);
        if (kXf) mbar_init(&a_ready_bar[s], kPair ? 8 : 4);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) { tmem_alloc_2sm(tmem_ptr_smem, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  int kblocks_total = 0;
  for (int s = 0; s < p.nseg; ++s) kblocks_total += p.seg_taps[s] * p.seg_chunks[s];

  // kXf: every warpgroup sets its register budget at the top of its own role branch (ptxas allocates per region; a budget set in
  // code that all roles share afterwards would cap every role at the smallest one)
  if (warp < 4) {
  if (kXf) setmaxnreg_ctl();
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (kHalo) {
      // activation rows: per (segment, 64-channel chunk) the rows y0-1 .. y0+2 (3x3) or y0, y0+1 (1x1) of the work item
      if (!(p.dbg & 1)) {
        int slot = 0;
        uint32_t ph = 0;
        for (int work = work0; work < p.total_work; work += work_stride) {
          const TileCoord t = decode_tile(p, work_to_tile(p, work, rank, 0));
          for (int s = 0; s < p.nseg; ++s) {
            const int nrows = p.seg_taps[s] == 9 ? 4 : 2;
            const int ybase = p.seg_taps[s] == 9 ? t.y0 - 1 : t.y0;
            for (int ch = 0; ch < p.seg_chunks[s]; ++ch) {
              for (int r = 0; r < nrows; ++r) {
                mbar_wait(&a_empty_bar[slot], ph ^ 1);
                if (!elect_one_sync()) {
                } else if (kPair && !kXf) {
                  const uint32_t fb = leader_addr(&a_full_bar[slot]);
                  mbar_arrive_expect_tx_cluster(fb, (uint32_t)kHaloRowBytes);
                  tma_load_4d_2sm(smem + slot * kHaloSlot, &p.mapA[s], fb, ch * kBlockK, t.x0 - 1, ybase + r, t.n0);
                } else {
                  mbar_arrive_expect_tx(&a_full_bar[slot], (uint32_t)kHaloRowBytes);
                  tma_load_4d(smem + slot * kHaloSlot, &p.mapA[s], &a_full_bar[slot], ch * kBlockK, t.x0 - 1, ybase + r, t.n0);
                }
                __syncwarp();
                if (++slot == nslots) { slot = 0; ph ^= 1; }
              }
            }
          }
        }
      }
    } else {
      // the whole warp walks the loop (converged); one elected lane issues the TMA instructions
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)stage_bytes;
      for (int work = work0; work < p.total_work; work += work_stride) {
        const TileCoord t = decode_tile(p, work_to_tile(p, work, rank, 0));
        const TileCoord t1 = decode_tile(p, work_to_tile(p, work, rank, mt - 1));
        for (int s = 0; s < p.nseg; ++s) {
          const int taps = p.seg_taps[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = (taps == 9) ? (tap / 3 - 1) : 0;
            const int dx = (taps == 9) ? (tap % 3 - 1) : 0;
            for (int ch = 0; ch < p.seg_chunks[s]; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_dst = smem + stage * stage_bytes;
              uint8_t* b_dst = a_dst + a_bytes;
              if (!elect_one_sync()) {
                // not the issuing lane
              } else if (kPair) {
                // this CTA's pixel tile + its half of the weight rows; completion is counted on the LEADER's barrier
                const uint32_t fb = leader_addr(&full_bar[stage]);
                mbar_arrive_expect_tx_cluster(fb, tx_bytes);
                tma_load_4d_2sm(a_dst, &p.mapA[s], fb, ch * kBlockK, t.x0 + dx, t.y0 + dy, t.n0);
                if (mt == 2) tma_load_4d_2sm(a_dst + kABytes, &p.mapA[s], fb, ch * kBlockK, t1.x0 + dx, t1.y0 + dy, t1.n0);
                tma_load_2d_2sm(b_dst, &p.mapB[s], fb, ch * kBlockK, tap * p.Cout_pad + t.nn0 + rank * p.b_rows);
              } else {
                mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
                tma_load_4d(a_dst, &p.mapA[s], &full_bar[stage], ch * kBlockK, t.x0 + dx, t.y0 + dy, t.n0);
                if (mt == 2) tma_load_4d(a_dst + kABytes, &p.mapA[s], &full_bar[stage], ch * kBlockK, t1.x0 + dx, t1.y0 + dy, t1.n0);
                tma_load_2d(b_dst, &p.mapB[s], &full_bar[stage], ch * kBlockK, tap * p.Cout_pad + t.nn0);
              }
              __syncwarp();
              if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (kHalo) {
      if (rank == 0) {     // converged warp, elected lane issues MMAs and commits
        int aslot = 0, bst = 0, acc = 0;
        uint32_t aph = 0, bph = 0, acc_phase = 0;
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(b_ring);
        // loop invariants of the plan, read once (left as p.x inside the elected block they were re-fetched from the constant bank
        // for every batch of eight MMAs)
        const uint32_t idesc = p.idesc, halo_bo = (uint32_t)p.halo_bo, bn_cols = (uint32_t)p.BN;
        const int n_stages = p.num_stages, use_ws = p.ws, dbg1 = p.dbg & 1;
        auto commit = [](uint64_t* bar) { if (kPair) umma_commit_2sm(bar); else umma_commit(bar); };   // pair: arrives in both CTAs
        for (int work = work0; work < p.total_work; work += work_stride) {
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)acc * 2u * bn_cols, d1 = d0 + bn_cols;
          uint32_t accum = 0;
          for (int s = 0; s < p.nseg; ++s) {
            const bool k3 = p.seg_taps[s] == 9;
            const int nrows = k3 ? 4 : 2;
            for (int ch = 0; ch < p.seg_chunks[s]; ++ch) {
              int rs[4];
              uint32_t rp[4];
              for (int r = 0; r < nrows; ++r) {
                rs[r] = aslot; rp[r] = aph;
                if (++aslot == nslots) { aslot = 0; aph ^= 1; }
              }
              const int ndy = k3 ? 3 : 1;
              for (int dyi = 0; dyi < ndy; ++dyi) {
                // tile 0 (row y0) reads slot row dyi, tile 1 (row y0+1) reads slot row dyi+1
                if (!dbg1) {
                  uint64_t* a_rdy = kXf ? a_ready_bar : a_full_bar;
                  if (dyi == 0) mbar_wait(&a_rdy[rs[0]], rp[0]);
                  mbar_wait(&a_rdy[rs[dyi + 1]], rp[dyi + 1]);
                }
                tc_fence_after();
                const int ndx = k3 ? 3 : 1;
                for (int dxi = 0; dxi < ndx; ++dxi) {
                  const uint32_t roff = k3 ? (uint32_t)dxi : 1u;       // rows into the slot: dx + 1
                  if (!dbg1) mbar_wait(&full_bar[bst], bph);
                  tc_fence_after();
                  const uint32_t bo = halo_bo ? roff : 0u;
                  const uint64_t a0_desc = umma_desc_sw128_bo(a_base + rs[dyi] * kHaloSlot + roff * 128u, bo);
                  const uint64_t a1_desc = umma_desc_sw128_bo(a_base + rs[dyi + 1] * kHaloSlot + roff * 128u, bo);
                  const uint64_t b_desc = umma_desc_sw128(b_base + bst * b_stage_bytes);
                  if (elect_one_sync()) {
#pragma unroll
                  for (int k = 0; k < kBlockK / 16; ++k) {
                    const uint32_t acc_k = (accum | (uint32_t)k) ? 1u : 0u;    // only the very first MMA of a work item overwrites
                    if (kPair) {
                      umma_bf16_ss_2sm(d0, a0_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, acc_k);
                      umma_bf16_ss_2sm(d1, a1_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, acc_k);
                    } else if (use_ws) {
                      umma_bf16_ss_ws_fill(d0, a0_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, acc_k);
                      umma_bf16_ss_ws_lastuse(d1, a1_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, acc_k);
                    } else {
                      umma_bf16_ss(d0, a0_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, acc_k);
                      umma_bf16_ss(d1, a1_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, acc_k);
                    }
                  }
                  commit(&empty_bar[bst]);
                  }
                  __syncwarp();
                  accum = 1;
                  if (++bst == n_stages) { bst = 0; bph ^= 1; }
                }
                // rows that no later tap of this chunk reads go back to the producer once the MMAs above retire
                if (elect_one_sync()) {
                  if (!k3) { commit(&a_empty_bar[rs[0]]); commit(&a_empty_bar[rs[1]]); }
                  else if (dyi < 2) commit(&a_empty_bar[rs[dyi]]);
                  else { commit(&a_empty_bar[rs[2]]); commit(&a_empty_bar[rs[3]]); }
                }
                __syncwarp();
              }
            }
          }
          if (elect_one_sync()) commit(&tmem_full_bar[acc]);
          __syncwarp();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    } else if (rank == 0) {
      // converged warp; the elected lane issues the MMAs and their commits (tcgen05.commit tracks the issuing thread)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t bn = (uint32_t)p.BN, idesc = p.idesc;
      const bool ws = p.ws != 0;
      for (int work = work0; work < p.total_work; work += work_stride) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * mt) * bn;
        for (int kb = 0; kb < kblocks_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint64_t a_desc = umma_desc_sw128(a_addr);
          const uint64_t a1_desc = umma_desc_sw128(a_addr + kABytes);
          const uint64_t b_desc = umma_desc_sw128(a_addr + a_bytes);
          if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128B swizzle atom: +2 in the (addr>>4) field
            const uint32_t accum = (kb | k) ? 1u : 0u;
            if (kPair) {
              umma_bf16_ss_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
              if (mt == 2) umma_bf16_ss_2sm(d_tmem + bn, a1_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
            } else if (mt == 2 && ws) {
              umma_bf16_ss_ws_fill(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
              umma_bf16_ss_ws_lastuse(d_tmem + bn, a1_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
            } else {
              umma_bf16_ss(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
              if (mt == 2) umma_bf16_ss(d_tmem + bn, a1_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
            }
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if (kPair) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (elect_one_sync()) {
          if (kPair) umma_commit_2sm(&tmem_full_bar[acc]); else umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================== halo mode: weight producer, one (chunk, tap) slab of BN x 64 per stage =====================
    if (kHalo && !(p.dbg & 1)) {
      int st = 0;
      uint32_t ph = 0;
      for (int work = work0; work < p.total_work; work += work_stride) {
        const TileCoord t = decode_tile(p, work_to_tile(p, work, rank, 0));
        for (int s = 0; s < p.nseg; ++s) {
          for (int ch = 0; ch < p.seg_chunks[s]; ++ch) {
            for (int tap = 0; tap < p.seg_taps[s]; ++tap) {
              mbar_wait(&empty_bar[st], ph ^ 1);
              if (!elect_one_sync()) {
              } else if (kPair) {   // this CTA's half of the weight rows
                const uint32_t fb = leader_addr(&full_bar[st]);
                mbar_arrive_expect_tx_cluster(fb, (uint32_t)b_stage_bytes);
                tma_load_2d_2sm(b_ring + st * b_stage_bytes, &p.mapB[s], fb, ch * kBlockK, tap * p.Cout_pad + t.nn0 + rank * p.b_rows);
              } else {
                mbar_arrive_expect_tx(&full_bar[st], (uint32_t)b_stage_bytes);
                tma_load_2d(b_ring + st * b_stage_bytes, &p.mapB[s], &full_bar[st], ch * kBlockK, tap * p.Cout_pad + t.nn0);
              }
              __syncwarp();
              if (++st == p.num_stages) { st = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== residual producer: TMA-loads the identity-skip tile of every output tile =====================
    if (p.res_tma && !(p.dbg & 2)) {     // converged warp, elected lane issues
      const int n_slabs = p.BN / 64;
      // nearest-up skip (unet.py:107,190-197): the 128-pixel output tile reads a (TH/2 x TW/2) box of the half-resolution source
      const int up = p.res_mode == 3 ? 1 : 0;
      const uint32_t slab_bytes = p.res_slab_bytes;
      uint32_t it = 0;
      for (int work = work0; work < p.total_work; work += work_stride) {
        for (int j = 0; j < mt; ++j) {
          const TileCoord t = decode_tile(p, work_to_tile(p, work, rank, j));
          for (int s0 = 0; s0 < n_slabs; s0 += 2, ++it) {
            const int ns = (n_slabs - s0) < 2 ? (n_slabs - s0) : 2;
            mbar_wait(res_empty_bar, (it & 1) ^ 1);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(res_full_bar, (uint32_t)ns * slab_bytes);
              for (int jj = 0; jj < ns; ++jj) {
                int c_off = t.nn0 + (s0 + jj) * 64;
                const CUtensorMap* mp = &p.mapRes;
                if (p.gn_red != nullptr && c_off >= p.gn_C0) { c_off -= p.gn_C0; mp = &p.mapRes2; }
                tma_load_4d(res_stage + jj * kSlabBytes, mp, res_full_bar, c_off, t.x0 >> up, t.y0 >> up, t.n0);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  }
  } else if (kXf && warp >= 12) {
    setmaxnreg_xf();
    // ===================== operand transform: act(A x + B) on the landed rows, in place =====================
    // Thread tt owns the logical 16-byte chunk j = tt & 7 (channels 8j .. 8j+7 of the 64-channel block) of the pixels (tt >> 3) + 16 i;
    // the 128B swizzle keeps it in physical chunk j ^ (pixel & 7), the same for all of them.  Pixels outside the image were
    // zero-filled by TMA and stay zero: the conv pads the NORMALISED activation (unet.py:185,211).
    if (!(p.dbg & 1)) {
      const int tt = (threadIdx.x - 384) & 127, wg = (threadIdx.x - 384) >> 7;
      const int j = tt & 7, pb = tt >> 3;
      uint32_t rowctr = 0;
      const uint32_t col = (uint32_t)((j ^ (pb & 7)) << 4);
      int slot = 0;
      uint32_t ph = 0;
      for (int work = work0; work < p.total_work; work += work_stride) {
        const TileCoord t = decode_tile(p, work_to_tile(p, work, rank, 0));
        const bool edge_l = t.x0 == 0, edge_r = t.x0 + 128 >= p.W;
        for (int s = 0; s < p.nseg; ++s) {
          const int nrows = p.seg_taps[s] == 9 ? 4 : 2;
          const int ybase = p.seg_taps[s] == 9 ? t.y0 - 1 : t.y0;
          const float* abp = p.xf_ab[s];
          for (int ch = 0; ch < p.seg_chunks[s]; ++ch) {
            float A[8], B[8];
            if (abp != nullptr) {
              const float4* q = reinterpret_cast<const float4*>(abp + ((size_t)t.n0 * p.xf_C + ch * kBlockK + j * 8) * 2);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 v = __ldg(q + e);
                A[2 * e] = v.x; B[2 * e] = v.y; A[2 * e + 1] = v.z; B[2 * e + 1] = v.w;
              }
              if (p.xf_silu) {   // silu(u) = h + h tanh(h), h = u / 2: halving (A, B) is exact, so h (and the result) keep silu_fast's bits
#pragma unroll
                for (int e = 0; e < 8; ++e) { A[e] *= 0.5f; B[e] *= 0.5f; }
              }
            }
            for (int r = 0; r < nrows; ++r, ++rowctr) {
              // Every warpgroup observes EVERY phase of every slot's barrier, also for the rows the other one transforms: with an
              // odd ring depth a slot alternates between the two warpgroups, a group that only waited for its own rows would
              // test a parity two phases old, and a late TMA of the row in between (rows land out of order under L2 misses)
              // let that wait pass on the previous contents of the slot - a double transform, a_ready counts off by four, and
              // eventually a pipeline that never completes (seen once in ~9000 evaluations as a trapped wait).
              const bool mine = (int)(rowctr & 1u) == wg;
              if (mine || !(p.dbg & 16)) mbar_wait(&a_full_bar[slot], ph);   // dbg 16: the old protocol (regression experiments only)
              if (!mine) {                         // the other warpgroup's row
                if (++slot == nslots) { slot = 0; ph ^= 1; }
                continue;
              }
              const int y = ybase + r;
              if (abp != nullptr && y >= 0 && y < p.H && !(p.dbg & 4)) {
                uint8_t* row = smem + slot * kHaloSlot + col;
                auto xform = [&](uint4& u) {
                  float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
                  f0.x = fmaf(A[0], f0.x, B[0]); f0.y = fmaf(A[1], f0.y, B[1]); f1.x = fmaf(A[2], f1.x, B[2]); f1.y = fmaf(A[3], f1.y, B[3]);
                  f2.x = fmaf(A[4], f2.x, B[4]); f2.y = fmaf(A[5], f2.y, B[5]); f3.x = fmaf(A[6], f3.x, B[6]); f3.y = fmaf(A[7], f3.y, B[7]);
                  if (p.dbg & 8) return;
                  if (p.xf_silu) {
                    f0.x = fmaf(f0.x, fast_tanh(f0.x), f0.x); f0.y = fmaf(f0.y, fast_tanh(f0.y), f0.y);
                    f1.x = fmaf(f1.x, fast_tanh(f1.x), f1.x); f1.y = fmaf(f1.y, fast_tanh(f1.y), f1.y);
                    f2.x = fmaf(f2.x, fast_tanh(f2.x), f2.x); f2.y = fmaf(f2.y, fast_tanh(f2.y), f2.y);
                    f3.x = fmaf(f3.x, fast_tanh(f3.x), f3.x); f3.y = fmaf(f3.y, fast_tanh(f3.y), f3.y);
                  }
                  u.x = pack_bf16(f0.x, f0.y); u.y = pack_bf16(f1.x, f1.y); u.z = pack_bf16(f2.x, f2.y); u.w = pack_bf16(f3.x, f3.y);
                };
#pragma unroll
                for (int i0 = 0; i0 < 8; i0 += 4) {
                  uint4 v[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<const uint4*>(row + (pb + 16 * (i0 + i)) * 128);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const int px = pb + 16 * (i0 + i);
                    if (!(edge_l && px == 0)) {       // pixel x0-1 outside the image
                      xform(v[i]);
                      *reinterpret_cast<uint4*>(row + px * 128) = v[i];
                    }
                  }
                }
                if (pb < 2 && !(edge_r && pb == 1)) {   // pixels 128 and 129 (= x0+128, outside the image on the right edge)
                  uint4 v = *reinterpret_cast<const uint4*>(row + (pb + 128) * 128);
                  xform(v);
                  *reinterpret_cast<uint4*>(row + (pb + 128) * 128) = v;
                }
                fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
              }
              __syncwarp();
              if (lane == 0) {
                if (kPair) mbar_arrive_cluster(leader_addr(&a_ready_bar[slot])); else mbar_arrive(&a_ready_bar[slot]);
              }
              if (++slot == nslots) { slot = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 12 && (kXf || p.tma_epilogue)) {
    if (kXf) setmaxnreg_epi();
    // ===================== epilogue (TMA store): 8 warps; warp (q, g) owns TMEM lanes [32q, 32q+32) x slab g =====================
    const int q = warp & 3;
    const int g = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 128;
    const int n_slabs = p.BN / 64;
    const uint32_t swz = (uint32_t)(row & 7);
    // row of this output pixel's skip value inside the TMA-loaded residual tile
    int res_r = row;
    if (p.res_mode == 3) {
      const int tw = row % p.TW, th = (row / p.TW) % p.TH, tn = row / (p.TW * p.TH);
      res_r = (tn * (p.TH >> 1) + (th >> 1)) * (p.TW >> 1) + (tw >> 1);
    }
    const uint32_t res_swz = (uint32_t)(res_r & 7);
    int acc = 0;
    uint32_t acc_phase = 0, res_it = 0;
    for (int work = work0; work < p.total_work; work += work_stride) {
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (p.dbg & 2) {
        tc_fence_before();
        if (kPair) mbar_arrive_cluster(leader_addr(&tmem_empty_bar[acc])); else mbar_arrive(&tmem_empty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
     for (int mj = 0; mj < mt; ++mj) {
      const TileCoord t = decode_tile(p, work_to_tile(p, work, rank, mj));
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * mt + mj) * p.BN);
      for (int s0 = 0; s0 < n_slabs; s0 += 2) {
        const bool active = (s0 + g) < n_slabs;
        uint32_t v0[32], v1[32];
        if (active) {
          tmem_ld_32x32(t_row + (uint32_t)((s0 + g) * 64), v0);
          tmem_ld_32x32(t_row + (uint32_t)((s0 + g) * 64 + 32), v1);
          tmem_ld_wait();
        }
        if (s0 + 2 >= n_slabs && mj == mt - 1) {   // last TMEM read of this work item: hand the accumulator back to the (leader's) MMA warp
          tc_fence_before();
          if (kPair) mbar_arrive_cluster(leader_addr(&tmem_empty_bar[acc]));
          else mbar_arrive(&tmem_empty_bar[acc]);
        }
        // the staging tile (and bias_s) may be rewritten once the previous TMA store has finished reading it
        // TMA stores are issued (and their bulk groups waited on) by one elected lane of the first epilogue warp, from converged code
        if (warp == 4) {
          if (elect_one_sync()) tma_store_wait_read();
          __syncwarp();
        }
        if (p.bias != nullptr && epi_tid < 128 && s0 * 64 + epi_tid < p.BN) bias_s[epi_tid] = __ldg(p.bias + t.nn0 + s0 * 64 + epi_tid);
        named_bar_sync(1, kEpiThreads);
        if (p.res_tma) {
          mbar_wait(res_full_bar, res_it & 1);
          ++res_it;
        }
        if (active) {
          uint8_t* dst_row = staging + g * kSlabBytes + row * 128;
          const uint8_t* res_row = res_stage + g * kSlabBytes + res_r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {     // 8 chunks of 8 channels (16 B)
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(j < 4 ? v0[j * 8 + e] : v1[(j - 4) * 8 + e]);
            if (p.bias != nullptr) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + g * 64 + j * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + g * 64 + j * 8 + 4);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            const uint32_t off = ((uint32_t)j ^ swz) << 4;
            if (p.res_tma && p.gn_red == nullptr) {
              const uint4 ru = *reinterpret_cast<const uint4*>(res_row + (((uint32_t)j ^ res_swz) << 4));
              const float2 r0 = unpack_bf16(ru.x), r1 = unpack_bf16(ru.y), r2 = unpack_bf16(ru.z), r3 = unpack_bf16(ru.w);
              f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y; f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
            }
            uint4 u;
            u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
            *reinterpret_cast<uint4*>(dst_row + off) = u;
          }
          fence_proxy_async();   // make the staging writes visible to the TMA (async proxy)
        }
        if (p.res_tma && p.gn_red == nullptr) mbar_arrive(res_empty_bar);
        named_bar_sync(2, kEpiThreads);
        if (p.gn_red != nullptr) {
          // GroupNorm backward, pass 1, on the staged gradient tile g and the TMA-loaded x tile (same swizzled layout):
          // red[n][c] += (sum g_u, sum g_u x), g_u = g act'(A x + B).  Same conflict-free column mapping as the statistics below.
          const int wi = warp - 4, gs = wi >> 2, oct = wi & 3;
          if (s0 + gs < n_slabs) {
            const int j = oct * 8 + (lane & 7), rq = lane >> 3;
            const uint8_t* gbase = staging + gs * kSlabBytes + (j & 3) * 4;
            const uint8_t* xbase = res_stage + gs * kSlabBytes + (j & 3) * 4;
            const int ipi = 16 / p.TN;
            const int ch = t.nn0 + (s0 + gs) * 64 + 2 * j;
            for (int img = 0; img < p.TN; ++img) {
              const int n = t.n0 + img;
              const float4 abv = __ldg(reinterpret_cast<const float4*>(p.gn_ab + ((size_t)(n < p.N ? n : p.N - 1) * p.Cout + ch) * 2));
              float r10 = 0.f, r11 = 0.f, r20 = 0.f, r21 = 0.f;
              for (int i = img * ipi; i < (img + 1) * ipi; ++i) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                  const int r = 2 * rq + 8 * i + b;
                  const uint32_t off = (uint32_t)(r * 128) + ((((uint32_t)j >> 2) ^ (uint32_t)(r & 7)) << 4);
                  const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(gbase + off));
                  const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(xbase + off));
                  float g0 = g.x, g1 = g.y;
                  if (p.gn_silu) {
                    g0 *= dsilu_fast(fmaf(abv.x, x.x, abv.y));
                    g1 *= dsilu_fast(fmaf(abv.z, x.y, abv.w));
                  }
                  r10 += g0; r20 = fmaf(g0, x.x, r20);
                  r11 += g1; r21 = fmaf(g1, x.y, r21);
                }
              }
#pragma unroll
              for (int o = 8; o <= 16; o <<= 1) {
                r10 += __shfl_xor_sync(0xffffffffu, r10, o); r11 += __shfl_xor_sync(0xffffffffu, r11, o);
                r20 += __shfl_xor_sync(0xffffffffu, r20, o); r21 += __shfl_xor_sync(0xffffffffu, r21, o);
              }
              if (lane < 8 && n < p.N) {
                float* rd = p.gn_red + ((size_t)n * p.Cout + ch) * 2;
                atomicAdd(rd, r10); atomicAdd(rd + 1, r20); atomicAdd(rd + 2, r11); atomicAdd(rd + 3, r21);
              }
            }
          }
          mbar_arrive(res_empty_bar);   // every epilogue thread, after its reads of the x tile
        }
        if (p.chan_stats != nullptr) {
          // Per-(image, channel) sum / sum of squares of the STORED bf16 values for the next GroupNorm (nn.py:17-19), read back
          // from the staged tile.  Warp (gs, oct) owns 8 channel pairs of slab gs; lane = rq*8 + pair, rq picks rows
          // 2rq + 8i + {0,1}: the four row classes of one LDS hit four distinct swizzle chunk pairs -> conflict-free.
          const int wi = warp - 4, gs = wi >> 2, oct = wi & 3;
          if (s0 + gs < n_slabs) {
            const int j = oct * 8 + (lane & 7), rq = lane >> 3;
            const uint8_t* sbase = staging + gs * kSlabBytes + (j & 3) * 4;
            const int ipi = 16 / p.TN;   // 8-row groups per image
            const int ch = t.nn0 + (s0 + gs) * 64 + 2 * j;
            for (int img = 0; img < p.TN; ++img) {
              float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
              for (int i = img * ipi; i < (img + 1) * ipi; ++i) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                  const int r = 2 * rq + 8 * i + b;
                  const uint32_t wv = *reinterpret_cast<const uint32_t*>(sbase + r * 128 + ((((uint32_t)j >> 2) ^ (uint32_t)(r & 7)) << 4));
                  const float2 v = unpack_bf16(wv);
                  a0 += v.x; a1 += v.y; q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
                }
              }
#pragma unroll
              for (int o = 8; o <= 16; o <<= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                q0 += __shfl_xor_sync(0xffffffffu, q0, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
              }
              if (lane < 8 && t.n0 + img < p.N) {
                float* st = p.chan_stats + ((size_t)(t.n0 + img) * p.Cout + ch) * 2;
                atomicAdd(st, a0); atomicAdd(st + 1, q0); atomicAdd(st + 2, a1); atomicAdd(st + 3, q1);
              }
            }
          }
        }
        if (warp == 4) {
          if (elect_one_sync()) {
            const int ns = (n_slabs - s0) < 2 ? (n_slabs - s0) : 2;
            for (int j = 0; j < ns; ++j)
              tma_store_4d(staging + j * kSlabBytes, &p.mapOut, t.nn0 + (s0 + j) * 64, t.x0, t.y0, t.n0);
            tma_store_commit();
          }
          __syncwarp();
        }
      }
     }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (warp == 4) {
      if (elect_one_sync()) tma_store_wait_all();
      __syncwarp();
    }
  } else if (!kXf && warp >= 4 && warp < 8) {
    // ===================== legacy epilogue: 4 warps, warp q owns TMEM lanes [32q, 32q+32) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const size_t HW = (size_t)p.H * p.W;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile);
      const int n = t.n0 + tn, y = t.y0 + th, x = t.x0 + tw;
      const bool valid = (n < p.N) && (y < p.H) && (x < p.W);
      const size_t pix = ((size_t)n * p.H + y) * p.W + x;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
      const int ncols = p.BN < 32 ? p.BN : 32;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        if (p.BN >= 32) {
          tmem_ld_32x32(t_row + (uint32_t)c0, v);
        } else {
          uint32_t v16[16];
          tmem_ld_32x16(t_row + (uint32_t)c0, v16);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v16[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0;
        }
        tmem_ld_wait();
        const int col0 = t.nn0 + c0;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) f[j] += __ldg(p.bias + col0 + j);
        }
        if (valid && p.res_mode != 0) {
          // skip path of the ResBlock / attention residual (bf16 NHWC with Cout channels)
          const int C = p.Cout;
          if (p.res_mode == 1) {
            const __nv_bfloat16* r = p.residual + pix * C + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float r8[8];
              load8_bf16(r + g * 8, r8);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[g * 8 + j] += r8[j];
            }
          } else if (p.res_mode == 2) {
            // 2x2 average pool of the [N,2H,2W,C] source (Downsample, unet.py:136)
            const int W2 = p.W * 2;
            const __nv_bfloat16* r = p.residual + (((size_t)n * (p.H * 2) + 2 * y) * W2 + 2 * x) * C + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float a8[8], b8[8], c8[8], d8[8];
              load8_bf16(r + g * 8, a8);
              load8_bf16(r + C + g * 8, b8);
              load8_bf16(r + (size_t)W2 * C + g * 8, c8);
              load8_bf16(r + (size_t)W2 * C + C + g * 8, d8);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[g * 8 + j] += 0.25f * ((a8[j] + b8[j]) + (c8[j] + d8[j]));
            }
          } else {
            // nearest-neighbour x2 of the [N,H/2,W/2,C] source (Upsample, unet.py:107)
            const int Wh = p.W / 2;
            const __nv_bfloat16* r = p.residual + (((size_t)n * (p.H / 2) + (y >> 1)) * Wh + (x >> 1)) * C + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float r8[8];
              load8_bf16(r + g * 8, r8);
#pragma unroll
              for (int j = 0; j < 8; ++j) f[g * 8 + j] += r8[j];
            }
          }
        }
        if (valid) {
          if (p.out_mode == 0) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * p.Cout + col0;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = pack_bf16(f[g * 8 + 0], f[g * 8 + 1]);
              u.y = pack_bf16(f[g * 8 + 2], f[g * 8 + 3]);
              u.z = pack_bf16(f[g * 8 + 4], f[g * 8 + 5]);
              u.w = pack_bf16(f[g * 8 + 6], f[g * 8 + 7]);
              *reinterpret_cast<uint4*>(o + g * 8) = u;
            }
          } else if (p.out_mode == 2) {
            // fp32 NHWC rows of Cout_pad values (tap-folded head / first-layer input-gradient GEMMs, gathered by tap_gather)
            float* o = reinterpret_cast<float*>(p.out) + pix * p.Cout_pad + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              if (g * 4 < ncols)
                *reinterpret_cast<float4*>(o + g * 4) = make_float4(f[g * 4] * p.out_scale, f[g * 4 + 1] * p.out_scale, f[g * 4 + 2] * p.out_scale, f[g * 4 + 3] * p.out_scale);
          } else {
            float* o = reinterpret_cast<float*>(p.out) + ((size_t)n * p.Cout) * HW + (size_t)y * p.W + x;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.Cout) o[(size_t)(col0 + j) * HW] = f[j] * p.out_scale;
          }
        }
        if (p.chan_stats != nullptr) {
          // per-(image, channel) sum / sum of squares of the stored value, for the next GroupNorm (nn.py:17-19).
          // Rows of one warp may span two images only when TN > 1 (8x8 level): reduce per image.
          const int n_lo = __shfl_sync(0xffffffffu, n, 0);
          const int n_hi = __shfl_sync(0xffffffffu, n, 31);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float val = valid ? f[j] : 0.f;
            val = __bfloat162float(__float2bfloat16(val));  // statistics of what is stored
            if (n_lo == n_hi) {
              float s1 = warp_sum(val), s2 = warp_sum(val * val);
              if (lane == 0 && n_lo < p.N) {
                atomicAdd(p.chan_stats + ((size_t)n_lo * p.Cout + col0 + j) * 2, s1);
                atomicAdd(p.chan_stats + ((size_t)n_lo * p.Cout + col0 + j) * 2 + 1, s2);
              }
            } else if (valid) {
              atomicAdd(p.chan_stats + ((size_t)n * p.Cout + col0 + j) * 2, val);
              atomicAdd(p.chan_stats + ((size_t)n * p.Cout + col0 + j) * 2 + 1, val * val);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer may still be reading this CTA's operands / arriving on its barriers
  if (warp == 2) {
    tc_fence_after();
    if (kPair) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct ConvPlan {
  ConvParams params;
  int grid;
  int xf;          // 512-thread variant with the operand-transform warps
  size_t smem_bytes;
};

// N tile: the widest of {256, 128, 64} dividing Cout_pad that still yields about one work item per SM (deep 8x8 / 16x16
// levels have few pixel tiles; narrower N tiles keep more SMs busy there); 16 / 32 for the tiny-Cout convs.
static int pick_bn(int cout_pad, int m_tiles) {
  if (cout_pad == 16 || cout_pad == 32) return cout_pad;
  if (cout_pad % 64 != 0) return -1;
  const int sms = num_sms();
  int best = -1;
  for (int bn : {256, 128, 64}) {
    if (cout_pad % bn != 0) continue;
    if (best < 0) best = bn;
    if ((long)m_tiles * (cout_pad / bn) >= sms) return bn;
    best = bn;   // not enough tiles yet: try narrower
  }
  return best;
}

static void set_smem_attr() {
  static bool attr_set = false;
  if (attr_set) return;
  cudaFuncSetAttribute(conv_gemm_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(conv_gemm_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(conv_gemm_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(conv_gemm_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(conv_gemm_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(conv_gemm_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  // host-mapped trap record (mbar_wait_line); failing to get one only loses the diagnostics
  if (cudaHostAlloc(reinterpret_cast<void**>(&g_trap_host), 64 * sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess) {
    memset(g_trap_host, 0, 64 * sizeof(unsigned int));
    unsigned int* dptr = nullptr;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), g_trap_host, 0) == cudaSuccess)
      cudaMemcpyToSymbol(g_trap_rec, &dptr, sizeof(dptr));
  } else {
    g_trap_host = nullptr;
  }
  cudaGetLastError();
  attr_set = true;
}

static int max_active_pairs(size_t smem_bytes, bool xf) {
  set_smem_attr();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * num_sms());
  cfg.blockDim = dim3(xf ? kXfThreads : kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  const cudaError_t e = xf ? cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<true, true, true>, &cfg)
                           : cudaOccupancyMaxActiveClusters(&n, conv_gemm_kernel<true, false, false>, &cfg);
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

// pixel-tile geometry of a conv (shared by the plan builder and the fusion predicates)
static void tile_dims(int H, int W, int* TW, int* TH) {
  *TW = W >= 16 ? 16 : W;
  const int rem = *TW > 0 ? kBlockM / *TW : 0;
  *TH = H >= rem ? rem : H;
}
// Can the GroupNorm-backward reduction of this conv's output ride in its epilogue (kdip_conv_desc.gn_*)?
bool conv_can_fuse_gn_reduce(const kdip_conv_desc* d) {
  if (getenv("KDIP_UNFUSED_GNRED") != nullptr) return false;
  if (d->out_mode != 0 || d->res_mode != 0 || d->Cout != d->Cout_pad || d->Cout % 64 != 0) return false;
  if (d->gn_C0 <= 0 || d->gn_C0 > d->Cout || d->gn_C0 % 64 != 0) return false;
  if (d->gn_C0 < d->Cout && d->gn_x1 == nullptr) return false;
  int TW, TH;
  tile_dims(d->H, d->W, &TW, &TH);
  if (TW <= 0 || kBlockM % TW != 0 || TH <= 0 || (kBlockM / TW) % TH != 0) return false;
  return d->W % TW == 0 && d->H % TH == 0;
}

// halo pipeline: a 3x3 first segment on an image at least 128 pixels wide, bf16 NHWC output in 64-channel slabs
static bool conv_uses_halo(const kdip_conv_desc* d) {
  if (!(d->seg[0].taps == 9 && d->W % 128 == 0 && d->H % 2 == 0 && d->out_mode == 0 && d->Cout == d->Cout_pad && d->Cout % 64 == 0 &&
        d->res_mode != 2))
    return false;
  // KDIP_CONV_HALO=0 selects the 8x16-tile pipeline everywhere (A/B measurements)
  if (getenv("KDIP_CONV_HALO") && atoi(getenv("KDIP_CONV_HALO")) == 0) return false;
  if (const char* e = getenv("KDIP_CONV_TMAEPI")) { if (atoi(e) == 0) return false; }
  return true;
}
// Can the GroupNorm apply (+SiLU) of this conv's input ride on its operand path (kdip_conv_desc.in_ab)?  Only the halo pipeline
// has the transform warps; KDIP_FUSE_GNAPPLY=0 keeps the separate gn_apply pass everywhere (A/B measurements).
bool conv_can_fuse_gn_apply(const kdip_conv_desc* d) {
  if (getenv("KDIP_FUSE_GNAPPLY") && atoi(getenv("KDIP_FUSE_GNAPPLY")) == 0) return false;
  return conv_uses_halo(d);
}

int conv_plan_build(const kdip_conv_desc* d, ConvPlan* plan) {
  KDIP_REQUIRE(d != nullptr && plan != nullptr, KDIP_EINVAL, "conv: null descriptor");
  KDIP_REQUIRE(d->nseg >= 1 && d->nseg <= 3, KDIP_EINVAL, "conv: nseg must be 1..3 (got %d)", d->nseg);
  KDIP_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, KDIP_ESHAPE, "conv: bad geometry N=%d H=%d W=%d", d->N, d->H, d->W);
  ConvParams& p = plan->params;
  memset(&p, 0, sizeof(p));
  int BN = 0;
  KDIP_REQUIRE(d->Cout >= 1 && d->Cout <= d->Cout_pad, KDIP_ESHAPE, "conv: Cout=%d > Cout_pad=%d", d->Cout, d->Cout_pad);
  KDIP_REQUIRE(d->out_mode >= 0 && d->out_mode <= 2, KDIP_EINVAL, "conv: out_mode must be 0, 1 or 2 (got %d)", d->out_mode);
  KDIP_REQUIRE(d->out_mode != 2 || (d->Cout_pad % 16 == 0 && d->Cout_pad <= 256 && d->res_mode == 0), KDIP_ESHAPE,
               "conv: fp32 NHWC output needs Cout_pad a multiple of 16 (<= 256) and no residual");
  KDIP_REQUIRE(d->out_mode != 0 || (d->Cout == d->Cout_pad && d->Cout % 32 == 0), KDIP_ESHAPE,
               "conv: bf16 NHWC output needs Cout == Cout_pad, multiple of 32 (got %d/%d)", d->Cout, d->Cout_pad);
  KDIP_REQUIRE(d->res_mode == 0 || (d->residual != nullptr && d->out_mode == 0), KDIP_EINVAL,
               "conv: residual needs a pointer and bf16 output");
  KDIP_REQUIRE(d->res_mode != 3 || (d->H % 2 == 0 && d->W % 2 == 0), KDIP_ESHAPE, "conv: nearest-up residual needs even H,W");
  KDIP_REQUIRE(d->out != nullptr && ((uintptr_t)d->out % 16) == 0, KDIP_EALIGN, "conv: out must be 16B aligned");
  KDIP_REQUIRE(d->chan_stats == nullptr || d->out_mode == 0, KDIP_EINVAL, "conv: chan_stats only with bf16 output");

  p.N = d->N; p.H = d->H; p.W = d->W;
  p.halo = conv_uses_halo(d) ? 1 : 0;
  // Measured on B200 under sustained load (tools/time_unet.py 32 50), AFTER the issue loops became converged-warp + elect.sync:
  // 8x16 tiles 41.0 ms per UNet evaluation, halo pipeline 40.7 ms, halo pipeline as CTA pairs 38.1 ms.  (With the old lane-0 issue
  // loops, which paced every variant at ~1000 clk per k-block, the same A/B read 47.4 / 48.0 / 49.0 ms and the halo was opt-in.)
  // KDIP_CONV_HALO=0 selects the 8x16-tile pipeline everywhere.
  bool xf = false;
  for (int s = 0; s < d->nseg; ++s) {
    p.xf_ab[s] = d->in_ab[s];
    if (d->in_ab[s] != nullptr) xf = true;
  }
  p.xf_C = d->in_ab_C; p.xf_silu = d->in_silu;

  if (xf) {
    KDIP_REQUIRE(p.halo, KDIP_ESHAPE, "conv: the fused GroupNorm apply (in_ab) needs the halo pipeline: 3x3 first segment, W a multiple of 128, even H, bf16 NHWC output in 64-channel slabs");
    KDIP_REQUIRE(d->in_ab_C > 0 && d->in_ab_C % 8 == 0, KDIP_EINVAL, "conv: in_ab_C=%d must be the (positive, multiple of 8) channel count of the (A, B) table", d->in_ab_C);
    for (int s = 0; s < d->nseg; ++s)
      KDIP_REQUIRE(((uintptr_t)d->in_ab[s] % 16) == 0, KDIP_EALIGN, "conv: in_ab[%d] must be 16B aligned", s);
  }
  plan->xf = xf ? 1 : 0;
  // Measured on B200 (tools/halo_probe.py): the 128B swizzle of tcgen05.mma operands is a function of the ABSOLUTE shared-memory
  // address, so a descriptor may start any number of 128-byte rows into a 1024-byte atom with base offset 0; setting the
  // matrix-base-offset field to the row phase gives wrong products.  KDIP_HALO_BASEOFF=1 re-enables it for that probe only.
  p.dbg = getenv("KDIP_CONV_DBG") ? atoi(getenv("KDIP_CONV_DBG")) : 0;
  p.ws = getenv("KDIP_CONV_WS") ? atoi(getenv("KDIP_CONV_WS")) : 1;   // weight-stationary pairs: 0.2-2 % faster, bit-identical results
  p.halo_bo = 0;
  if (const char* e = getenv("KDIP_HALO_BASEOFF")) p.halo_bo = atoi(e) ? 1 : 0;
  p.TW = p.halo ? 128 : (d->W >= 16 ? 16 : d->W);
  KDIP_REQUIRE(kBlockM % p.TW == 0, KDIP_ESHAPE, "conv: W=%d must be >=16 or a power of two", d->W);
  int rem = kBlockM / p.TW;
  p.TH = d->H >= rem ? rem : d->H;
  KDIP_REQUIRE(rem % p.TH == 0, KDIP_ESHAPE, "conv: H=%d must be >=%d or a power of two", d->H, rem);
  p.TN = rem / p.TH;
  p.tiles_x = (d->W + p.TW - 1) / p.TW;
  p.tiles_y = (d->H + p.TH - 1) / p.TH;
  p.tiles_n = (d->N + p.TN - 1) / p.TN;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  BN = p.halo ? (d->Cout_pad % 128 == 0 ? 128 : 64) : pick_bn(d->Cout_pad, m_tiles);   // halo: two accumulators per buffer, BN <= 128
  // 8x16-tile pipeline: N tile 128 with two pixel tiles per work item as CTA pairs ("128p1m2") measured best or tied on every
  // narrow-level shape with at least ~32 pair work items (tools/gpu_round38.sh sweep, B=32: 16^2 512->512 46.7 -> 34.9 us,
  // 16^2 1024->512 80.2 -> 56.3, 64^2 128->128 48.1 -> 38.8, 32^2 768->256 95.6 -> 82.2); the 8^2 level keeps N tile 64 pairs.
  bool force_p1m2 = false;
  if (!p.halo && d->Cout_pad % 128 == 0 && m_tiles % 4 == 0 && (long)(m_tiles / 4) * (d->Cout_pad / 128) >= 32 && d->out_mode == 0 &&
      !getenv("KDIP_CONV_BN") && !getenv("KDIP_CONV_MT") && !getenv("KDIP_CONV_PAIR") && !getenv("KDIP_CONV_PAIRMT")) {
    BN = 128;
    force_p1m2 = true;
  }
  if (const char* e = getenv("KDIP_CONV_BN")) {   // tuning override for the 8x16-tile pipeline
    const int v = atoi(e);
    if (!p.halo && (v == 64 || v == 128 || v == 256) && d->Cout_pad % v == 0) BN = v;
  }
  KDIP_REQUIRE(BN > 0, KDIP_ESHAPE, "conv: Cout_pad=%d unsupported (need 16, 32 or a multiple of 64)", d->Cout_pad);
  p.BN = BN;
  p.n_tiles = d->Cout_pad / BN;
  p.total_tiles = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles;
  p.Cout_pad = d->Cout_pad;
  p.Cout = d->Cout;
  p.nseg = d->nseg;
  // TMA-store epilogue for bf16 NHWC outputs in 64-channel slabs without pooled / upsampled skips or fused statistics
  // (fused statistics read whole staged tiles: only when the pixel tiles divide the image)
  const bool whole_tiles = (d->W % p.TW == 0) && (d->H % p.TH == 0);
  // nearest-up skips ride the TMA path when every tile maps to a whole half-resolution box
  const bool up_ok = d->res_mode == 3 && whole_tiles && p.TW % 2 == 0 && (p.TH % 2 == 0 || p.halo);
  p.tma_epilogue = (d->out_mode == 0 && BN >= 64 && d->Cout == d->Cout_pad && (d->res_mode == 0 || d->res_mode == 1 || up_ok) &&
                    (d->chan_stats == nullptr || whole_tiles)) ? 1 : 0;
  p.res_tma = (p.tma_epilogue && d->res_mode != 0) ? 1 : 0;
  if (d->gn_red != nullptr) {
    KDIP_REQUIRE(conv_can_fuse_gn_reduce(d) && p.tma_epilogue, KDIP_ESHAPE,
                 "conv: fused GroupNorm-backward reduction needs bf16 NHWC output, no residual, 64-channel source splits and whole pixel tiles");
    KDIP_REQUIRE(d->gn_x0 != nullptr && d->gn_ab != nullptr && ((uintptr_t)d->gn_x0 % 16) == 0 && ((uintptr_t)d->gn_x1 % 16) == 0, KDIP_EINVAL,
                 "conv: fused GroupNorm-backward reduction needs 16B-aligned sources and the (A, B) table");
    p.res_tma = 1;
  }
  // CTA pairs whenever the pixel tiles pair up
  p.pair = (p.tma_epilogue && (m_tiles % 2 == 0)) ? 1 : 0;
  // tuning switches for A/B measurements (tools/time_unet.py): KDIP_CONV_PAIR=0, KDIP_CONV_TMAEPI=0
  if (const char* e = getenv("KDIP_CONV_TMAEPI")) { if (atoi(e) == 0 && d->gn_red == nullptr) { p.tma_epilogue = 0; p.pair = 0; p.res_tma = 0; } }
  if (const char* e = getenv("KDIP_CONV_PAIR")) { if (atoi(e) == 0) p.pair = 0; }
  // Two pixel tiles per work item when the N tile is narrow: the SM's L2 read port (~64 B/clk) feeds M128 x N128 x K64 MMAs
  // (256 clk) with 32 KB per k-block = 128 B/clk, i.e. at most half rate; sharing the weight stage between two accumulators
  // needs 48 KB per 512 clk.  TMEM: 2 buffers x 2 tiles x BN columns <= 512.  Only when the wave quantisation stays benign.
  p.mt = 1;
  if (p.tma_epilogue && BN <= 128 && m_tiles % 2 == 0) {
    const int sms = num_sms();
    const long w1 = (long)m_tiles * p.n_tiles, w2 = w1 / 2;
    const double e1 = (double)w1 / (double)(((w1 + sms - 1) / sms) * sms), e2 = (double)w2 / (double)(((w2 + sms - 1) / sms) * sms);
    if (w2 >= sms && e2 >= e1 - 0.08) p.mt = 2;
  }
  if (const char* e = getenv("KDIP_CONV_MT")) {   // 1: never; 2: whenever legal (tests force it on small shapes)
    if (atoi(e) == 1) p.mt = 1;
    if (atoi(e) == 2 && p.tma_epilogue && BN <= 128 && m_tiles % 2 == 0) p.mt = 2;
  }
  // mt = 2 as CTA pairs (four pixel tiles per work item): 40.3 vs 41.0 ms per UNet evaluation (sustained); KDIP_CONV_PAIRMT=0: single CTAs
  if (p.mt == 2 && (m_tiles % 4 != 0 || (getenv("KDIP_CONV_PAIRMT") && atoi(getenv("KDIP_CONV_PAIRMT")) == 0))) p.pair = 0;
  if (force_p1m2 && p.tma_epilogue) { p.mt = 2; p.pair = 1; }
  if (p.halo) {
    KDIP_REQUIRE(p.tma_epilogue, KDIP_EINVAL, "conv: internal error, halo pipeline without the TMA epilogue");
    // CTA pairs (M = 256 MMAs, weight rows split across the two CTAs) keep the per-SM shared-memory traffic of the N = 128 MMAs
    // under 128 B/clk: operand reads 96 B/clk + TMA writes ~30 B/clk, against 128 + 46 for single CTAs (measured bound ~75 %)
    // Measured (sustained, whole UNet): pairs 38.1 ms vs single CTAs 40.7 ms.  KDIP_HALO_PAIR=0 selects single CTAs.
    p.mt = 2;
    p.pair = (m_tiles % 4 == 0) ? 1 : 0;
    if (const char* e = getenv("KDIP_HALO_PAIR")) { if (atoi(e) == 0) p.pair = 0; }
  }
  p.b_rows = p.pair ? BN / 2 : BN;
  p.total_work = (m_tiles / (p.mt * (p.pair ? 2 : 1))) * p.n_tiles;
  p.idesc = umma_idesc_bf16(p.pair ? 2 * kBlockM : kBlockM, BN);
  p.bias = d->bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  p.res_mode = d->res_mode;
  p.out = d->out;
  p.out_mode = d->out_mode;
  p.out_scale = d->out_mode != 0 ? d->out_scale : 1.0f;
  p.chan_stats = d->chan_stats;
  p.gn_ab = d->gn_ab; p.gn_red = d->gn_red; p.gn_C0 = d->gn_C0; p.gn_silu = d->gn_silu;
  p.gn_two = (d->gn_red != nullptr && d->gn_x1 != nullptr && d->gn_C0 < d->Cout) ? 1 : 0;

  for (int s = 0; s < d->nseg; ++s) {
    const kdip_conv_seg& sg = d->seg[s];
    KDIP_REQUIRE(sg.taps == 9 || sg.taps == 1, KDIP_EINVAL, "conv: taps must be 9 or 1 (got %d)", sg.taps);
    KDIP_REQUIRE(sg.C > 0 && sg.C % kBlockK == 0, KDIP_ESHAPE, "conv: segment channels %d not a multiple of 64", sg.C);
    KDIP_REQUIRE(sg.act != nullptr && sg.wgt != nullptr, KDIP_EINVAL, "conv: null segment pointer");
    KDIP_REQUIRE(((uintptr_t)sg.act % 16) == 0 && ((uintptr_t)sg.wgt % 16) == 0, KDIP_EALIGN, "conv: segment pointers must be 16B aligned");
    p.seg_taps[s] = sg.taps;
    p.seg_chunks[s] = sg.C / kBlockK;
    int rc = p.halo ? encode_tmap_bf16_4d(&p.mapA[s], sg.act, (uint64_t)sg.C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N, kBlockK,
                                          (uint32_t)kHaloPix, 1u, 1u)
                    : encode_tmap_bf16_4d(&p.mapA[s], sg.act, (uint64_t)sg.C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N, kBlockK,
                                          (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN);
    if (rc != KDIP_OK) return rc;
    rc = encode_tmap_bf16_2d(&p.mapB[s], sg.wgt, (uint64_t)sg.C, (uint64_t)sg.taps * d->Cout_pad, kBlockK, (uint32_t)p.b_rows);
    if (rc != KDIP_OK) return rc;
  }

  int extra = 1024 /*bias + barriers*/;
  if (p.tma_epilogue) {
    extra += kStagingBytes + (p.res_tma ? kStagingBytes : 0);
    int rc = encode_tmap_bf16_4d(&p.mapOut, d->out, (uint64_t)d->Cout, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N, 64,
                                 (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN);
    if (rc != KDIP_OK) return rc;
    if (d->gn_red != nullptr) {
      rc = encode_tmap_bf16_4d(&p.mapRes, d->gn_x0, (uint64_t)d->gn_C0, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N, 64,
                               (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN);
      if (rc != KDIP_OK) return rc;
      if (p.gn_two) {
        rc = encode_tmap_bf16_4d(&p.mapRes2, d->gn_x1, (uint64_t)(d->Cout - d->gn_C0), (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N, 64,
                                 (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN);
        if (rc != KDIP_OK) return rc;
      }
    } else if (p.res_tma) {
      KDIP_REQUIRE(((uintptr_t)d->residual % 16) == 0, KDIP_EALIGN, "conv: residual must be 16B aligned");
      const int sh = d->res_mode == 3 ? 1 : 0;
      const uint32_t bh = (p.TH >> sh) > 0 ? (uint32_t)(p.TH >> sh) : 1u;     // halo tiles are one row high
      rc = encode_tmap_bf16_4d(&p.mapRes, d->residual, (uint64_t)d->Cout, (uint64_t)(d->W >> sh), (uint64_t)(d->H >> sh), (uint64_t)d->N, 64,
                               (uint32_t)(p.TW >> sh), bh, (uint32_t)p.TN);
      if (rc != KDIP_OK) return rc;
      p.res_slab_bytes = (uint32_t)(p.TW >> sh) * bh * (uint32_t)p.TN * 128u;
    }
    if (d->gn_red != nullptr) p.res_slab_bytes = kSlabBytes;
  }
  const int stage_bytes = p.halo ? p.b_rows * 128 : p.mt * kABytes + p.b_rows * 128;
  // the transform stage adds latency between a row's landing and its first MMA: one more slot of look-ahead (measured: UNet forward
  // 13.48 -> 13.36 ms at B=32; the 128-pixel level 157 -> 126 us per 128->128 conv); the plain pipeline gains nothing from it
  p.halo_slots = xf ? kHaloSlots + 1 : kHaloSlots;
  if (const char* e = getenv("KDIP_HALO_SLOTS")) { const int v = atoi(e); if (v >= 5 && v <= kMaxHaloSlots) p.halo_slots = v; }
  const int fixed = p.halo ? p.halo_slots * kHaloSlot : 0;
  const int budget = 227 * 1024 - extra - 1024 /*align slack*/ - 512 - fixed;
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* e = getenv("KDIP_CONV_STAGES")) { const int v = atoi(e); if (v >= 2 && v < stages) stages = v; }
  KDIP_REQUIRE(stages >= 2, KDIP_ESHAPE, "conv: not enough shared memory for a 2-stage pipeline (BN=%d)", BN);
  p.num_stages = stages;
  plan->smem_bytes = (size_t)fixed + (size_t)stages * stage_bytes + 1024 /*align slack*/ + extra + 512;
  if (getenv("KDIP_CONV_DEBUG"))
    fprintf(stderr, "[kdip conv] N=%d H=%d W=%d Cout=%d K=%d taps=%d nseg=%d BN=%d stages=%d pair=%d mt=%d halo=%d tmaepi=%d res=%d work=%d smem=%zu\n", d->N, d->H,
            d->W, d->Cout, d->seg[0].C, d->seg[0].taps, d->nseg, BN, stages, p.pair, p.mt, p.halo, p.tma_epilogue, d->res_mode, p.total_work, plan->smem_bytes);
  int sms = num_sms();
  if (p.pair) {
    // a persistent kernel must be fully co-resident: ask the driver how many CTA pairs fit at this shared-memory size
    int clusters = max_active_pairs(plan->smem_bytes, xf);
    if (clusters <= 0 || clusters > sms / 2) clusters = sms / 2;
    if (getenv("KDIP_CONV_DEBUG")) fprintf(stderr, "[kdip conv] pair clusters co-resident: %d\n", clusters);
    plan->grid = 2 * (p.total_work < clusters ? p.total_work : clusters);
  } else {
    plan->grid = p.total_work < sms ? p.total_work : sms;
  }
  return KDIP_OK;
}

int conv_plan_launch(const ConvPlan* plan, cudaStream_t stream) {
  set_smem_attr();
  if (plan->params.pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(plan->grid);
    cfg.blockDim = dim3(plan->xf ? kXfThreads : kThreads);
    cfg.dynamicSmemBytes = plan->smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    if (plan->xf) KDIP_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, true, true>, plan->params));
    else if (plan->params.halo) KDIP_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, true, false>, plan->params));
    else KDIP_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, false, false>, plan->params));
    count_launch();
    return KDIP_OK;
  }
  if (plan->xf) conv_gemm_kernel<false, true, true><<<plan->grid, kXfThreads, plan->smem_bytes, stream>>>(plan->params);
  else if (plan->params.halo) conv_gemm_kernel<false, true, false><<<plan->grid, kThreads, plan->smem_bytes, stream>>>(plan->params);
  else conv_gemm_kernel<false, false, false><<<plan->grid, kThreads, plan->smem_bytes, stream>>>(plan->params);
  KDIP_LAUNCH_CHECK();
  return KDIP_OK;
}

ConvPlan* conv_plan_new() { return new ConvPlan(); }
void conv_plan_free(ConvPlan* p) { delete p; }

}  // namespace kdip

struct kdip_conv_plan {
  kdip::ConvPlan plan;
};

extern "C" int kdip_conv_plan_create(const kdip_conv_desc* d, kdip_conv_plan** out) {
  KDIP_REQUIRE(out != nullptr, KDIP_EINVAL, "conv_plan_create: null out");
  kdip_conv_plan* p = new kdip_conv_plan();
  int rc = kdip::conv_plan_build(d, &p->plan);
  if (rc != KDIP_OK) {
    delete p;
    return rc;
  }
  *out = p;
  return KDIP_OK;
}
extern "C" int kdip_conv_plan_run(const kdip_conv_plan* p, kdip_stream_t s) {
  KDIP_REQUIRE(p != nullptr, KDIP_EINVAL, "conv_plan_run: null plan");
  return kdip::conv_plan_launch(&p->plan, (cudaStream_t)s);
}
extern "C" void kdip_conv_plan_destroy(kdip_conv_plan* p) { delete p; }
extern "C" int kdip_conv_trap_read(unsigned int* out, int n) {
  if (kdip::g_trap_host == nullptr || out == nullptr) return 0;
  const int cnt = (int)kdip::g_trap_host[0];
  for (int i = 0; i < n && i < 64; ++i) out[i] = kdip::g_trap_host[i];
  return cnt;
}
