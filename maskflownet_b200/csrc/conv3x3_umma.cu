// conv3x3_umma.cu -- the 3x3 convolutions of the decoder, the context network and the feature pyramid (SURVEY.md section
// 8f, row N2) on the 5th-generation tensor cores: tcgen05.mma with the accumulators in tensor memory (TMEM).  Contract as in
// conv3x3.cu (channel slices of a level buffer in, bias + LeakyReLU'ed channel slice out, fp32-accurate through the bf16
// hi/lo split: three MMAs hi*lo + lo*hi + hi*hi per product); reference call sites network/MaskFlownet.py:147-300.
//
// Implicit GEMM without im2col:   D[pixel, f] = sum_{tap, c} X[pixel + tap offset, c] * Wt[tap][c][f]
//   * M tile = 128 consecutive output pixels of one image row; a work tile is R = 2 rows x 128 pixels x all output channels
//     (two 128 x N fp32 accumulators in TMEM).  PERSISTENT kernel, one CTA per SM, tiles dealt round-robin; the
//     accumulators are double-buffered in TMEM (2 x 2N <= 512 columns for N <= 128), so tile i's epilogue runs under tile
//     i+1's MMAs, and the producers / weight loader simply run ahead across tile boundaries.
//   * K is walked as (16-channel chunk) x (tap).  Per chunk the producer warps convert the input rows the nine taps touch
//     from fp32 NCHW into split bf16 in the *no-swizzle K-major core-matrix layout*: plane [8-channel group][pixel] with
//     16 bytes per entry.  In that layout a tap shift is nothing but a different start address (+16 bytes per pixel), so
//     all nine taps are nine shared-memory descriptors over ONE converted tile: (start, LBO = plane pitch, SBO = 128 B).
//     Stride 2 de-interleaves even / odd pixels so the same holds (see the geometry helpers).
//   * weights are pre-packed (mfn_conv3x3_pack_weights) into per-(chunk, tap) images of the same layout and streamed by
//     one thread with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) through a deep ring (up to 16 stages: the
//     L2 -> shared latency of a tile is several times its MMA time).
//   * one thread issues the MMAs (M=128, N=CoutP, K=16, kind::f16, bf16 x bf16 -> fp32); tcgen05.commit releases the
//     weight / input stages and signals the four epilogue warps, which read the accumulators with tcgen05.ld (lane =
//     pixel, 16 output channels per instruction), add the bias, apply LeakyReLU and store NCHW (coalesced 128 B per
//     plane row and warp) -- or, for the transposed convolutions, scatter 2x2 sub-pixel phases (depth-to-space).
#include "mma_tiles.cuh"

namespace mfn {
namespace um {
using c3::smem_u32;
using c3::split_pair;

constexpr int MT = 128;          // pixels per M tile
constexpr int R = 2;             // output rows per CTA
constexpr int NTHREADS = 512;    // warps 0, 2: MMA issuers (warp 0 owns TMEM), warp 1: weight loader, warps 4..7: epilogue, rest: producers
constexpr int NPROD = 9;         // producer warps: 3, 8..15
constexpr int NISSUE = R;        // MMA-issuing threads: one per output row (lane 0 of warps 0 and 2)
constexpr int MAX_AS = 4, MAX_WS = 24;
constexpr int BATCH = 4;         // producer items (32 entries x 8 channels) in flight per warp
constexpr int MAX_ACC = 8;       // accumulator sets in TMEM (narrow layers: 2 rows x 32 columns each -- a deep ring hides the
                                 // commit -> epilogue -> acc_empty round trip, which is longer than such a tile's MMAs)
constexpr int BAR_BYTES = 1024;  // a_full/a_empty [MAX_AS], w_full/w_empty [MAX_WS], acc_full/acc_empty [MAX_ACC], TMEM pointer

// Geometry of the converted input tile: `nslots` image rows of PW entries each.
//   stride 1 (any dilation d): rows y0 - d .. y0 + R - 1 + d (or the 3R rows the taps touch when d >= R); entry p of a row is
//     pixel x0 - d + p; tap (ky, kx) of output row r starts at entry slot(r, ky) * PW + kx * d.
//   stride 2 (d = 1): rows 2 y0 - 1 .. 2 y0 + 2R - 1; a row is de-interleaved into its even pixels 2 (x0 + p), p < 128, and
//     its odd pixels 2 (x0 - 1 + p - 128) + 1, p >= 128, so that "next output pixel" is again "next entry": tap kx reads
//     the odd block from 0 (kx = 0), the even block (kx = 1) or the odd block from 1 (kx = 2).
__host__ __device__ inline int n_slots(int stride, int dil) { return stride == 2 ? 2 * R + 1 : (dil >= R ? 3 * R : R + 2 * dil); }
__host__ __device__ inline int row_pitch(int stride, int dil) { return stride == 2 ? 2 * MT + 1 : MT + 2 * dil; }
__host__ __device__ inline int slot_of(int r, int ky, int stride, int dil) {
  return stride == 2 ? 2 * r + ky : (dil >= R ? ky * R + r : r + ky * dil);
}
__host__ __device__ inline int tap_xoff(int kx, int stride, int dil) {
  return stride == 2 ? (kx == 1 ? 0 : (kx == 0 ? MT : MT + 1)) : kx * dil;
}
__host__ __device__ inline int row_of_slot(int slot, int y0, int stride, int dil) {
  return stride == 2 ? 2 * y0 - 1 + slot : (dil >= R ? y0 + (slot % R) + (slot / R - 1) * dil : y0 - dil + slot);
}
// ext = 2 ("band" mode of K3 through linearity, warp_lin.cu): the input is a VIRTUAL image of (n + 6) rows / columns --
// the n real ones followed by [0, 0, first, 0, 0, last] -- so that ONE convolution also yields the 1-D convolutions of the
// first / last row and column and the four corner pixels that the MXNet-1.5 border rule needs.  Maps a virtual
// coordinate to the real one, or -1 (zero).
__host__ __device__ inline int band_map(int v, int n) { return v < n ? v : (v == n + 2 ? 0 : (v == n + 5 ? n - 1 : -1)); }
__host__ __device__ inline int x_of_entry(int p, int x0, int stride, int dil) {
  return stride == 2 ? (p < MT ? 2 * (x0 + p) : 2 * (x0 - 1 + p - MT) + 1) : x0 - dil + p;
}
// output channels padded to the MMA's N granularity
__host__ __device__ inline int cout_pad(int cout) { return cout <= 16 ? 16 : (cout + 15) / 16 * 16; }
// Narrow layers (N <= 64) are bound by the MMA *issue* rate (measured ~50 cycles per tcgen05.mma from one thread, whatever
// N: profiles/r01_ubench_tcgen05_mma_issue.txt), so they fold the hi / lo weight images into ONE operand of 2N rows:
//   D[:, 0:2N] += A_hi x [B_hi ; B_lo]      (N' = 2N)        D[:, 0:N] += A_lo x B_hi
// two instructions per product instead of three; the epilogue adds the two column blocks.
// Folding the long wide layers too (65 <= N <= 128, >= 16 chunks: one accumulator set of 2 x 2N = 512 TMEM columns,
// 100 instead of 120 B/clk of shared-memory operand reads) was measured neutral (579 -> 128: 0.685 -> 0.667 ms, 259 -> 128:
// 0.330 -> 0.349 ms; step 7.09 -> 7.07 ms) and is not used; the <true, 1> instance stays compiled for the experiment.
__host__ __device__ inline bool fold_hi_lo(int CoutP, int nChunks) {
  (void)nChunks;
  return CoutP <= 64;
}
// Taps per weight-ring stage (compile-time variants of the kernel).  Both rings are bound by their ROUND-TRIP latency
// (commit -> mbarrier -> waiting thread wakes -> copy / conversion -> mbarrier -> issuer wakes: ~2 us measured), not by
// bandwidth: a ring of S stages delivers S stages per round trip.  Narrow layers, whose MMAs per tap are short, therefore
// move 9 / 3 taps per bulk copy.
__host__ __device__ inline int taps_per_stage(int CoutP) { return CoutP <= 32 ? 9 : (CoutP <= 64 ? 3 : 1); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
// Bounded wait (2^28 polls, each of which suspends for the hardware's try_wait window): a protocol bug fails the launch
// with a trap instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: 8-row x 16-byte core matrices; SBO = distance between 8-row
// groups (M/N direction), LBO = distance between the two 8-element K groups of one MMA (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// Split-K plan (host: plan_split): work items [0, from) are whole tiles, the tiles [from, numTiles) are cut into k parts
// over the channel chunks.  Their raw sums go to ws[part][n - n_lo][Cout][y - y_lo (rh rows)][OW].
struct SplitK {
  int k, from, n_lo, y_lo, rh;
  long long part_stride;
  float* ws;
  // tile -> (n, ty, tx) without integer division: ceil(2^32 / (tilesX tilesY)) and ceil(2^32 / tilesX), or 0 when the
  // products could overflow the exactness bound (then the kernel divides)
  uint32_t magic_tp, magic_tx;
  int as_wide;   // input stages of the wide (single-tap weight stage) layers: 3 or 2 (smem_map)
};
__device__ __forceinline__ void decode_tile(int tile, int tilesX, int tilesY, const SplitK& sk, int& tx, int& ty, int& n) {
  if (sk.magic_tp) {
    n = (int)__umulhi((uint32_t)tile, sk.magic_tp);
    const int rem = tile - n * tilesX * tilesY;
    ty = sk.magic_tx ? (int)__umulhi((uint32_t)rem, sk.magic_tx) : rem;   // magic_tx == 0: tilesX == 1
    tx = rem - ty * tilesX;
  } else {
    tx = tile % tilesX;
    ty = (tile / tilesX) % tilesY;
    n = tile / (tilesX * tilesY);
  }
}
struct Work {
  int tile, part, cb, ce;   // part < 0: a whole tile
};
__device__ __forceinline__ Work decode_work(int w, const SplitK& sk, int nChunks) {
  Work r;
  if (sk.k <= 1 || w < sk.from) {
    r.tile = w; r.part = -1; r.cb = 0; r.ce = nChunks;
  } else {
    const int u = w - sk.from, t = u / sk.k;
    r.tile = sk.from + t;
    r.part = u - t * sk.k;
    r.cb = r.part * nChunks / sk.k;
    r.ce = (r.part + 1) * nChunks / sk.k;
  }
  return r;
}

struct SmemMap {
  int a_lo, a_stage, w_tile, w_stage, w_off, bar_off, total, AS, WS;
};
// stage counts from the shared-memory budget: 3 input stages when that still leaves >= 8 weight stages, else 2
__host__ __device__ inline SmemMap smem_map(int E, int CoutP, int as_wide = 3) {
  SmemMap m;
  m.a_lo = 2 * E * 16;              // hi image: two 8-channel planes of E entries
  m.a_stage = 2 * m.a_lo;           // hi + lo
  m.w_tile = 64 * CoutP;            // [hi | lo][2 planes][CoutP][16 B]
  const int budget = 227 * 1024 - BAR_BYTES;
  m.w_stage = taps_per_stage(CoutP) * m.w_tile;
  // input stages: narrow layers (several taps per weight stage) take 4 when >= 4 weight stages still fit; wide layers keep
  // the weight ring deep (their weight stages are single taps) and take 3
  if (taps_per_stage(CoutP) > 1)
    m.AS = (4 * m.a_stage + 4 * m.w_stage <= budget) ? 4 : ((3 * m.a_stage + 3 * m.w_stage <= budget) ? 3 : 2);
  else   // as_wide (tuning "conv_as"): 2 trades an input stage for four more single-tap weight stages
    m.AS = (as_wide >= 3 && 3 * m.a_stage + 8 * m.w_stage <= budget) ? 3 : 2;
  int ws = (budget - m.AS * m.a_stage) / m.w_stage;
  m.WS = ws > MAX_WS ? MAX_WS : ws;
  m.w_off = m.AS * m.a_stage;
  m.bar_off = m.w_off + m.WS * m.w_stage;
  m.total = m.bar_off + BAR_BYTES;
  return m;
}
}  // namespace um

// UMMA weight image: [16-channel chunk c][tap][hi | lo][8-channel plane kc][f (CoutP)][8 x bf16 = 16 bytes]
// (narrow layers: [chunk][tap][plane][hi f.. | lo f..][16 bytes], see fold_hi_lo)
__global__ void conv3x3_pack_umma_kernel(const float* __restrict__ w, unsigned char* __restrict__ packed, int Cin, int Cout,
                                         int CoutP, int nChunks16) {
  const long long total = (long long)nChunks16 * 9 * CoutP * 8;   // (c, tap, f, channel pair)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);                     // channel pair inside the chunk
    const int f = (int)((i >> 3) % CoutP);
    const int tap = (int)((i / (8LL * CoutP)) % 9);
    const int c = (int)(i / (8LL * CoutP * 9));
    const int ch = 16 * c + 2 * j;
    float a = 0.f, b = 0.f;
    if (f < Cout) {
      if (ch < Cin) a = w[((size_t)f * Cin + ch) * 9 + tap];
      if (ch + 1 < Cin) b = w[((size_t)f * Cin + ch + 1) * 9 + tap];
    }
    uint32_t hi, lo;
    c3::split_pair(a, b, hi, lo);
    const int wt = 64 * CoutP;
    unsigned char* tile = packed + ((size_t)c * 9 + tap) * wt;
    if (um::fold_hi_lo(CoutP, nChunks16)) {   // [8-channel plane][hi rows | lo rows][16 B]
      const int off = (j >> 2) * (2 * CoutP * 16) + f * 16 + (j & 3) * 4;
      *reinterpret_cast<uint32_t*>(tile + off) = hi;
      *reinterpret_cast<uint32_t*>(tile + CoutP * 16 + off) = lo;
    } else {                       // [hi | lo][8-channel plane][rows][16 B]
      const int off = (j >> 2) * (CoutP * 16) + f * 16 + (j & 3) * 4;
      *reinterpret_cast<uint32_t*>(tile + off) = hi;
      *reinterpret_cast<uint32_t*>(tile + wt / 2 + off) = lo;
    }
  }
}

template <bool FOLD, int TPS>
__global__ void __launch_bounds__(um::NTHREADS, 1)
    conv3x3_umma_kernel(const float* __restrict__ x, long long x_bs, const unsigned char* __restrict__ wpack,
                        const float* __restrict__ bias_arg, float* __restrict__ out_base, long long out_bs, int Cin, int H, int W,
                        int OH, int OW, int Cout, int CoutP, int nChunks, float slope_arg, int tilesX, int tilesY, int numTiles,
                        int stride, int dil, int out_mode_arg, int tmem_cols, int nacc, int ext, um::SplitK sk) {
  using namespace um;
  // out_mode_arg = mode | (linear_prefix << 8): the first linear_prefix output channels are written WITHOUT the activation
  // (a second, linear head sharing the input pass of an activated layer: network.py folds pred_flow / pred_mask over the
  // dense block's input into its last convolution)
  // Split-K (sk.k > 1): `numTiles` counts WORK ITEMS.  Items below sk.from are whole tiles; the tiles from sk.from on are
  // cut into sk.k parts over the channel chunks -- part p walks chunks [p nChunks / k, (p + 1) nChunks / k) and writes its
  // RAW partial sums (no bias, no activation) to the workspace; conv3x3_umma_reduce_kernel finishes that region.
  const int out_mode_k = out_mode_arg & 0xff, lin_prefix_k = out_mode_arg >> 8;
  extern __shared__ __align__(128) unsigned char smem[];
  const int nslots = n_slots(stride, dil), PW = row_pitch(stride, dil), E = nslots * PW;
  const SmemMap sm = smem_map(E, CoutP, sk.as_wide);
  const int AS = sm.AS, WS = sm.WS;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + sm.bar_off;
  const uint32_t a_full = bar0, a_empty = bar0 + 8 * MAX_AS, w_full = bar0 + 16 * MAX_AS, w_empty = w_full + 8 * MAX_WS,
                 acc_full = w_empty + 8 * MAX_WS, acc_empty = acc_full + 8 * MAX_ACC, tmem_ptr = acc_empty + 8 * MAX_ACC;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + sm.bar_off + (tmem_ptr - bar0));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t plane = (size_t)H * W;

  if (tid == 0) {
    for (int i = 0; i < AS; ++i) {
      mbar_init(a_full + 8 * i, NPROD);
      mbar_init(a_empty + 8 * i, NISSUE);
    }
    for (int i = 0; i < WS; ++i) {
      mbar_init(w_full + 8 * i, 1);
      mbar_init(w_empty + 8 * i, NISSUE);
    }
    for (int i = 0; i < MAX_ACC; ++i) {
      mbar_init(acc_full + 8 * i, NISSUE);
      mbar_init(acc_empty + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // TMEM allocation: one warp, address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 || warp == 2) {
    // ============================ MMA issuers: lane 0 of warp 0 (output row 0) and of warp 2 (row 1) ============================
    // A single thread sustains one tcgen05.mma per ~50-100 cycles whatever N (profiles/r01_ubench_tcgen05_mma_issue.txt),
    // below the tensor pipe's 64 cycles at N = 128 once waits and commits are added -- so each output row (its own
    // accumulator, no ordering between the two) gets its own issuing thread; every stage barrier collects both commits.
    if (lane == 0) {
      const int r = warp >> 1;
      // instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), K-major both, N >> 3 @17, M >> 4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CoutP >> 3) << 17) | ((uint32_t)(MT >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CoutP >> 2) << 17) | ((uint32_t)(MT >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)E * 16u, b_lbo = (uint32_t)(FOLD ? 2 * CoutP : CoutP) * 16u;
      const uint32_t row_cols = (uint32_t)(FOLD ? 2 * CoutP : CoutP);   // TMEM columns per output row
      // ring positions / parities are running counters (no divisions); descriptor low words (address >> 4) of the nine
      // taps live in registers (the tap loop is fully unrolled), the high words are constants
      uint32_t a_off[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
        a_off[tap] = (uint32_t)((slot_of(r, tap / 3, stride, dil) * PW + tap_xoff(tap % 3, stride, dil)) * 16) >> 4;
      const uint64_t desc_hi_a = ((uint64_t)((a_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
      const uint64_t desc_hi_b = ((uint64_t)((b_lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
      uint32_t as = 0, aph = 0, ws = 0, wph = 0, acc = 0, accph = 0, j = 0;
      const uint32_t a_lo16 = (uint32_t)sm.a_lo >> 4, a_stage16 = (uint32_t)sm.a_stage >> 4, s_base16 = s_base >> 4;
      const uint32_t w_base16 = (s_base + (uint32_t)sm.w_off) >> 4, w_tile16 = (uint32_t)sm.w_tile >> 4, w_half16 = w_tile16 >> 1;
      const uint32_t w_stage16 = (uint32_t)sm.w_stage >> 4;
      for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x, ++j) {
        if (j >= (uint32_t)nacc) mbar_wait(acc_empty + 8 * acc, accph ^ 1);   // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (acc * (uint32_t)R + (uint32_t)r) * row_cols;
        const Work wk = decode_work(tile, sk, nChunks);
        const int cb = wk.cb, ce = wk.ce;
        for (int c = cb; c < ce; ++c) {
          mbar_wait(a_full + 8 * as, aph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_st16 = s_base16 + as * a_stage16;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (tap % TPS == 0) {
              mbar_wait(w_full + 8 * ws, wph);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t w16 = w_base16 + ws * w_stage16 + (uint32_t)(tap % TPS) * w_tile16;
            const uint64_t a_hi = desc_hi_a | (uint64_t)(a_st16 + a_off[tap]), a_lo = desc_hi_a | (uint64_t)(a_st16 + a_off[tap] + a_lo16);
            const uint64_t b_hi = desc_hi_b | (uint64_t)w16;
            if (FOLD) {
              umma_bf16(d, a_hi, b_hi, idesc2, (tap == 0 && c == cb) ? 0u : 1u);   // [hi*hi | hi*lo]
              umma_bf16(d, a_lo, b_hi, idesc, 1u);                                // += lo*hi into the first block
            } else {
              const uint64_t b_lo = desc_hi_b | (uint64_t)(w16 + w_half16);
              umma_bf16(d, a_hi, b_lo, idesc, (tap == 0 && c == cb) ? 0u : 1u);
              umma_bf16(d, a_lo, b_hi, idesc, 1u);
              umma_bf16(d, a_hi, b_hi, idesc, 1u);
            }
            if (tap % TPS == TPS - 1) {
              umma_commit(w_empty + 8 * ws);     // weight stage free once both rows' MMAs have read it
              if (++ws == (uint32_t)WS) { ws = 0; wph ^= 1; }
            }
          }
          umma_commit(a_empty + 8 * as);         // input stage free
          if (++as == (uint32_t)AS) { as = 0; aph ^= 1; }
        }
        umma_commit(acc_full + 8 * acc);         // this row's accumulator is complete
        if (++acc == (uint32_t)nacc) { acc = 0; accph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================ weight loader (one thread) ============================
    if (lane == 0) {
      uint32_t ws = 0, wph = 0;
      bool wrapped = false;
      for (int tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const Work wk = decode_work(tile, sk, nChunks);
        const int cb = wk.cb, ce = wk.ce;
        const unsigned char* src = wpack + (size_t)cb * 9 * sm.w_tile;
        for (int it = 9 * cb; it < 9 * ce; it += TPS, src += sm.w_stage) {
          if (wrapped) mbar_wait(w_empty + 8 * ws, wph ^ 1);
          mbar_arrive_expect_tx(w_full + 8 * ws, (uint32_t)sm.w_stage);
          bulk_g2s(s_base + (uint32_t)sm.w_off + ws * (uint32_t)sm.w_stage, src, (uint32_t)sm.w_stage, w_full + 8 * ws);
          if (++ws == (uint32_t)WS) { ws = 0; wph ^= 1; wrapped = true; }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================ epilogue (TMEM lanes 32*(warp&3) .. +31 = pixels of the M tile) ============================
    const int q = warp & 3;
    const uint32_t row_cols = (uint32_t)(FOLD ? 2 * CoutP : CoutP);
    uint32_t j = 0;
    for (int work = blockIdx.x; work < numTiles; work += gridDim.x, ++j) {
      const Work wk = decode_work(work, sk, nChunks);
      const int tile = wk.tile;
      int tx, ty, n;
      decode_tile(tile, tilesX, tilesY, sk, tx, ty, n);
      // a part of a split tile: raw sums into its slot of the workspace region [n_lo.., all channels, y_lo.., OW]
      const bool partial = wk.part >= 0;
      const float* const bias = partial ? nullptr : bias_arg;
      const float slope = partial ? 1.f : slope_arg;
      const int out_mode = partial ? 0 : out_mode_k, lin_prefix = partial ? 0 : lin_prefix_k;
      const size_t oplane0 = partial ? (size_t)sk.rh * OW : (size_t)OH * OW;    // plane pitch of plain NCHW output
      float* const out = partial ? sk.ws + (size_t)wk.part * (size_t)sk.part_stride + (size_t)(n - sk.n_lo) * Cout * oplane0 -
                                       (size_t)sk.y_lo * OW
                                 : out_base + (size_t)n * out_bs;
      const uint32_t acc = j % (uint32_t)nacc;
      mbar_wait(acc_full + 8 * acc, (j / (uint32_t)nacc) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int xx = tx * MT + 32 * q + lane;
#pragma unroll 1
      for (int r = 0; r < R; ++r) {
        const int y = ty * R + r;
        const bool okp = y < OH && xx < OW;
        const uint32_t t0 = tmem_base + ((uint32_t)(32 * q) << 16) + acc * (uint32_t)R * row_cols + (uint32_t)r * row_cols;
#pragma unroll 1
        for (int nc = 0; nc < CoutP / 16; ++nc) {
          // The epilogue is a straight line: the TMEM loads, then the 16 (independent, L1-resident) bias loads under their
          // latency, then arithmetic on all 16 channels, then predicated stores.  (Round 1 had one branch + one dependent
          // bias load per channel: ~80 cycles x 16 channels per row made the NARROW layers epilogue-bound -- ncu of the
          // 16 -> 16 pyramid layer: the issuers spun on acc_empty, the producers on a_empty.)
          uint32_t v[16];
          tmem_ld16(t0 + (uint32_t)(nc * 16), v);
          uint32_t v2[16];
          if (FOLD) tmem_ld16(t0 + (uint32_t)(CoutP + nc * 16), v2);   // second column block: the hi * lo term
          float bv[16];
          if (out_mode == 0) {
            if (bias != nullptr && nc * 16 + 16 <= Cout) {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) bv[jj] = __ldg(bias + nc * 16 + jj);
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                const int f = nc * 16 + jj;
                bv[jj] = (bias != nullptr && f < Cout) ? __ldg(bias + f) : 0.f;
              }
            }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (FOLD) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) v[jj] = __float_as_uint(__uint_as_float(v[jj]) + __uint_as_float(v2[jj]));
          }
          if (out_mode == 0) {
            const size_t oplane = oplane0;
            float* on = out + (size_t)(nc * 16) * oplane + (size_t)y * OW + xx;
            if (nc * 16 + 16 <= Cout && (lin_prefix <= nc * 16 || lin_prefix >= nc * 16 + 16)) {
              // whole group valid, one activation: a running pointer and one predicated store per channel
              const float sl = lin_prefix >= nc * 16 + 16 ? 1.f : slope;
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                const float t = __uint_as_float(v[jj]) + bv[jj];
                const float o = t > 0.f ? t : t * sl;
                if (okp) *on = o;
                on += oplane;
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                const int f = nc * 16 + jj;
                const float t = __uint_as_float(v[jj]) + bv[jj];
                const float o = leaky(t, f < lin_prefix ? 1.f : slope);
                if (f < Cout && okp) on[(size_t)jj * oplane] = o;
              }
            }
          } else {
            // depth-to-space: conv channel f' = (2 py + px) * F + f  ->  out[n][f][2y + py][2x + px], out is (F, 2 OH, 2 OW)
            const int F = Cout >> 2;
            const size_t oplane = (size_t)(2 * OH) * (2 * OW);
            float* on = out + (size_t)(2 * y) * (2 * OW) + 2 * xx;
            int ph = (nc * 16) / F, f = nc * 16 - ph * F;          // running (phase, channel) of conv channel nc * 16 + jj
            float bd[16];
            int fo[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const bool okc = nc * 16 + jj < Cout;
              bd[jj] = (bias != nullptr && okc) ? __ldg(bias + f) : 0.f;
              fo[jj] = okc ? (f << 2) | ph : -1;
              if (++f == F) { f = 0; ++ph; }
            }
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const float o = leaky(__uint_as_float(v[jj]) + bd[jj], slope);
              if (fo[jj] >= 0 && okp)
                on[(size_t)(fo[jj] >> 2) * oplane + (size_t)((fo[jj] >> 1) & 1) * (2 * OW) + (fo[jj] & 1)] = o;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * acc);   // TMEM reads of this accumulator are done (wait::ld above)
    }
  } else {
    // ============================ input producers (warps 3, 8..15) ============================
    // One instruction stream for every layer shape: the (tile, chunk, batch) space of this CTA is walked as ONE flat
    // sequence of batches (BATCH items of 32 entries x 8 channels per warp), software-pipelined over two register sets --
    // the loads of batch i+1 (possibly the next chunk, possibly the next TILE) are in flight while batch i is converted and
    // stored.  Round 1 pipelined only across the chunks of one tile: layers with one or two chunks (the pyramid's levels
    // 1-2) paid a full DRAM round trip per tile.  Geometry is arithmetic only (e / PW through a multiply-high).
    const int pw = warp < 4 ? 0 : warp - 7;   // warp 3 -> 0, warps 8..15 -> 1..8
    const int G = (E + 31) / 32;
    const bool one_plane = Cin <= 8;          // a single chunk whose channels 8..15 are zeros: plane 1 is cleared once, never loaded
    const int nItems = one_plane ? G : 2 * G;                        // item = (32 entries, 8-channel plane)
    const int nb = (nItems + NPROD * BATCH - 1) / (NPROD * BATCH);   // batches per stage (and warp)
    const uint32_t pw_magic = 0xFFFFFFFFu / (uint32_t)PW + 1u;       // e / PW == umulhi(e, magic)  (e * PW < 2^32)
    if (one_plane) {
      for (int st = 0; st < AS; ++st)
        for (int e = pw * 32 + lane; e < E; e += NPROD * 32) {
          unsigned char* d = smem + st * sm.a_stage + (E + e) * 16;
          *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(d + sm.a_lo) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    // load cursor (runs one batch ahead of the store cursor)
    int l_tile = blockIdx.x, l_c = 0, l_kb = 0, l_x0 = 0, l_y0 = 0;
    const float* l_xn = x;
    int l_ce = nChunks;        // end of this work item's chunk range (split-K)
    auto set_tile = [&]() {
      int tl = l_tile;
      if (sk.k > 1) {
        const Work wk = decode_work(l_tile, sk, nChunks);
        tl = wk.tile;
        l_c = wk.cb;
        l_ce = wk.ce;
      }
      int tx, ty, n;
      decode_tile(tl, tilesX, tilesY, sk, tx, ty, n);
      // ext = 1: "full" convolution -- the output grid is the input grid extended by one pixel on every side
      // (OH = H + 2, OW = W + 2; output (y, x) sits at input position (y - 1, x - 1)); used by K3 through linearity
      l_x0 = tx * MT - (ext ? 1 : 0);
      l_y0 = ty * R - (ext ? 1 : 0);
      l_xn = x + (size_t)n * x_bs;
    };
    auto advance = [&]() {   // false when this CTA's sequence is exhausted
      if (++l_kb < nb) return true;
      l_kb = 0;
      if (++l_c < l_ce) return true;
      l_c = 0;
      l_tile += gridDim.x;
      if (l_tile >= numTiles) return false;
      set_tile();
      return true;
    };
    // Geometry without branches: y = ymul * y0 + ya + yb * (slot >> 1) + yc * (slot & 1) + yd * slot, x likewise (see
    // row_of_slot / x_of_entry); addresses = one uniform 64-bit chunk base + 32-bit element offsets (16 planes < 2^32
    // elements, checked by the host), so a load costs one add and one IMAD.WIDE instead of a 64-bit add chain.
    const bool s2 = stride == 2, wide = !s2 && dil >= R;
    const int ymul = s2 ? 2 : 1, ya = s2 ? -1 : -dil, yb = wide ? dil : 0, yc = wide ? 1 : 0, yd = wide ? 0 : 1;
    const int xa = s2 ? 0 : -dil;
    const uint32_t planeu = (uint32_t)plane;
    auto load_batch = [&](float (&v)[BATCH][8]) {
      const float* xc = l_xn + (size_t)(16 * l_c) * plane;   // warp-uniform
      const int ybase = ymul * l_y0 + ya, xbase = ymul * l_x0 + xa;
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int t = pw + (l_kb * BATCH + b) * NPROD;
        const int kc = one_plane ? 0 : (t & 1);
        const int e = (one_plane ? t : (t >> 1)) * 32 + lane;
        const int slot = (int)__umulhi((uint32_t)e, pw_magic), pe = e - slot * PW;
        int y = ybase + yb * (slot >> 1) + yc * (slot & 1) + yd * slot;
        const int q = (s2 && pe >= MT) ? 1 : 0;              // stride 2: odd-pixel block of the de-interleaved row
        int xx = xbase + ymul * (pe - q * MT) - q;           // q = 1: 2 (x0 - 1 + pe - MT) + 1
        if (ext == 2) {
          y = y >= 0 ? band_map(y, H) : -1;
          xx = xx >= 0 ? band_map(xx, W) : -1;
        }
        const bool ok = t < nItems && e < E && (unsigned)y < (unsigned)H && (unsigned)xx < (unsigned)W;
        const int c0 = 16 * l_c + 8 * kc;
        uint32_t off = (uint32_t)(8 * kc) * planeu + (ok ? (uint32_t)(y * W + xx) : 0u);
        if (c0 + 8 <= Cin) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            v[b][jj] = ok ? __ldg(xc + off) : 0.f;
            off += planeu;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            v[b][jj] = (ok && c0 + jj < Cin) ? __ldg(xc + off) : 0.f;
            off += planeu;
          }
        }
      }
    };
    // store cursor: batch index inside the stage + the running position in the input ring
    uint32_t as = 0, aph = 0;
    bool wrapped = false;
    auto store_batch = [&](int kb, const float (&v)[BATCH][8]) {
      if (kb == 0 && wrapped) mbar_wait(a_empty + 8 * as, aph ^ 1);   // the MMAs that read this stage have completed
      unsigned char* a_st = smem + as * sm.a_stage;
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const int t = pw + (kb * BATCH + b) * NPROD;
        const int kc = one_plane ? 0 : (t & 1);
        const int e = (one_plane ? t : (t >> 1)) * 32 + lane;
        if (t < nItems && e < E) {
          uint4 hi, lo;
          split_pair(v[b][0], v[b][1], hi.x, lo.x);
          split_pair(v[b][2], v[b][3], hi.y, lo.y);
          split_pair(v[b][4], v[b][5], hi.z, lo.z);
          split_pair(v[b][6], v[b][7], hi.w, lo.w);
          unsigned char* dst = a_st + (kc * E + e) * 16;
          *reinterpret_cast<uint4*>(dst) = hi;
          *reinterpret_cast<uint4*>(dst + sm.a_lo) = lo;
        }
      }
      if (kb == nb - 1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + 8 * as);
        if (++as == (uint32_t)AS) { as = 0; aph ^= 1; wrapped = true; }
      }
    };
    float va[BATCH][8], vb[BATCH][8];
    set_tile();
    load_batch(va);
    int s_kb = 0;
    for (;;) {
      bool more = advance();
      if (more) load_batch(vb);
      store_batch(s_kb, va);
      if (!more) break;
      s_kb = l_kb;
      more = advance();
      if (more) load_batch(va);
      store_batch(s_kb, vb);
      if (!more) break;
      s_kb = l_kb;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
long long conv3x3_umma_packed_bytes(int Cin, int Cout) {
  return (long long)((Cin + 15) / 16) * 9 * 64 * um::cout_pad(Cout);
}

int conv3x3_umma_pack(const float* weight, unsigned char* packed, int Cin, int Cout, cudaStream_t st) {
  const int CoutP = um::cout_pad(Cout), nChunks16 = (Cin + 15) / 16;
  const long long total = (long long)nChunks16 * 9 * CoutP * 8;
  long long blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  conv3x3_pack_umma_kernel<<<(unsigned)blocks, 256, 0, st>>>(weight, packed, Cin, Cout, CoutP, nChunks16);
  return check_launch("conv3x3_pack_umma_kernel");
}

// Split-K second pass over the split region (samples n_lo.., rows y_lo..): out = act(sum_p parts[p] + bias), NCHW (with
// the linear prefix) or depth-to-space.
__global__ void conv3x3_umma_reduce_kernel(um::SplitK sk, int RN, const float* __restrict__ bias, float* __restrict__ out,
                                           long long out_bs, int Cout, int OH, int OW, float slope, int out_mode_arg) {
  const int out_mode = out_mode_arg & 0xff, lin_prefix = out_mode_arg >> 8;
  const long long total = (long long)RN * Cout * sk.rh * OW;
  const int F = Cout >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < sk.k; ++p) s += sk.ws[(size_t)p * (size_t)sk.part_stride + (size_t)i];
    const int x = (int)(i % OW), y = sk.y_lo + (int)((i / OW) % sk.rh), f = (int)((i / ((long long)OW * sk.rh)) % Cout);
    const int n = sk.n_lo + (int)(i / ((long long)OW * sk.rh * Cout));
    if (out_mode == 0) {
      const float b = bias ? __ldg(bias + f) : 0.f;
      out[(size_t)n * out_bs + ((size_t)f * OH + y) * OW + x] = leaky(s + b, f < lin_prefix ? 1.f : slope);
    } else {
      const int ph = f / F, ff = f - ph * F;
      const float b = bias ? __ldg(bias + ff) : 0.f;
      out[(size_t)n * out_bs + ((size_t)ff * (2 * OH) + (2 * y + (ph >> 1))) * (2 * OW) + 2 * x + (ph & 1)] = leaky(s + b, slope);
    }
  }
}

// Split-K plan.  Two cases, both about tiles being indivisible units of a persistent grid of 148 CTAs:
//   small images (levels 5-6: 2 x tiles <= SMs, up to 43 chunks walked serially per tile at ~2 us per chunk): every tile is
//     cut into k parts;
//   a short last round (level 2: 896 tiles = 6 x 148 + 8 -- the 8 left-over tiles cost a 7th round, 13.5 % of the layer):
//     only the tail tiles are cut, into as many parts as there are idle SMs, so the last round shrinks to 1/k of a tile.
//     The tail is kept inside the last sample and aligned to whole tile rows, so the split region is a row range.
// Returns k = 1 when nothing is split.
static um::SplitK plan_split(int N, int Cin, int H, int W, int Cout, int stride, int dil, int grid_cap) {
  using namespace um;
  (void)dil;
  SplitK sk = {1, 0, 0, 0, 0, 0, nullptr, 0u, 0u, 3};
  const int OH = (H - 1) / stride + 1, OW = (W - 1) / stride + 1, nChunks = (Cin + 15) / 16;
  const int tilesX = (OW + MT - 1) / MT, tilesY = (OH + R - 1) / R;
  const long long tiles = (long long)N * tilesX * tilesY;
  sk.from = (int)tiles;
  const int mode = tuning().conv_splitk;
  if (!mode || tiles >= (1 << 30)) return sk;
  const int sms = grid_cap > 0 && grid_cap < kNumSMs ? grid_cap : kNumSMs;
  const int kcap = mode > 1 ? mode : 32;
  if (2 * tiles <= sms) {                       // small image: split every tile
    int k = (int)(sms / tiles);
    if (k > nChunks / 3) k = nChunks / 3;       // at least 3 chunks per part
    if (k > 8) k = 8;
    if (k > kcap) k = kcap;
    if (k >= 2) {
      sk.k = k; sk.from = 0; sk.n_lo = 0; sk.y_lo = 0; sk.rh = OH;
    }
  } else if (tiles > sms) {                     // short last round: split the tail
    const long long rounds = tiles / sms;
    long long tail = tiles - rounds * sms;
    tail = (tail + tilesX - 1) / tilesX * tilesX;                 // whole tile rows
    int k = tail > 0 ? (int)(sms / tail) : 0;
    if (k > nChunks / 2) k = nChunks / 2;       // at least 2 chunks per part
    if (k > kcap) k = kcap;
    // measured (tools/conv_profile.py, level 2): pays for the long, tensor-bound layers (579 -> 128: 0.766 -> 0.685 ms,
    // 387 -> 96, 259 -> 128); layers with few chunks or Cout <= 64 lose the gain to the second launch
    if (tail > 0 && tail <= (long long)tilesX * tilesY && rounds <= 12 && k >= 2 && nChunks >= 16 && Cout > 64) {
      sk.k = k; sk.from = (int)(tiles - tail); sk.n_lo = N - 1;
      sk.y_lo = (int)((sk.from / tilesX) % tilesY) * R;
      sk.rh = OH - sk.y_lo;
    }
  }
  if (sk.k > 1) sk.part_stride = (long long)(N - sk.n_lo) * Cout * sk.rh * OW;
  return sk;
}

long long conv3x3_umma_workspace_bytes(int N, int Cin, int H, int W, int Cout, int stride, int dil) {
  const um::SplitK sk = plan_split(N, Cin, H, W, Cout, stride, dil, tuning().conv_grid_cap);
  return sk.k > 1 ? sk.k * sk.part_stride * 4 : 0;
}

// returns -1 when the shape does not fit this kernel (caller falls back to the mma.sync kernel)
int conv3x3_umma_launch(const float* x, long long x_bs, const unsigned char* wpack, const float* bias, float* out,
                        long long out_bs, int N, int Cin, int H, int W, int Cout, int stride, int dil, int out_mode,
                        float slope, cudaStream_t st, int ext, float* ws, long long ws_bytes) {
  using namespace um;
  if (Cout > 256 || (stride != 1 && !(stride == 2 && dil == 1))) return -1;
  if (ext != 0 && !((ext == 1 || ext == 2) && stride == 1 && dil == 1 && out_mode == 0)) return -1;
  if ((out_mode >> 8) != 0 && (out_mode & 0xff) != 0) return -1;   // linear prefix only with plain NCHW output
  const int CoutP = um::cout_pad(Cout), nChunks = (Cin + 15) / 16;
  const int E = n_slots(stride, dil) * row_pitch(stride, dil);
  const int as_wide = tuning().conv_as == 2 ? 2 : 3;
  const SmemMap sm = smem_map(E, CoutP, as_wide);
  if (sm.WS < 2 || E * 16 > 0x3FFF * 16) return -1;
  if ((long long)H * W >= (1LL << 27)) return -1;   // the producers address a 16-plane chunk with 32-bit element offsets
  const int grow = ext == 2 ? 8 : 2 * ext;   // ext 1: grid + 1 pixel per side; ext 2: + the six band rows / columns too
  const int OH = stride == 2 ? (H - 1) / 2 + 1 : H + grow, OW = stride == 2 ? (W - 1) / 2 + 1 : W + grow;
  static SmemOptIn opt0, opt1, opt3, opt9;
  {
    cudaError_t e = ensure_dyn_smem(conv3x3_umma_kernel<false, 1>, sm.total, opt0);
    if (e == cudaSuccess) e = ensure_dyn_smem(conv3x3_umma_kernel<true, 1>, sm.total, opt1);
    if (e == cudaSuccess) e = ensure_dyn_smem(conv3x3_umma_kernel<true, 3>, sm.total, opt3);
    if (e == cudaSuccess) e = ensure_dyn_smem(conv3x3_umma_kernel<true, 9>, sm.total, opt9);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(conv3x3_umma_kernel): %s", cudaGetErrorString(e));
  }
  const bool fold = fold_hi_lo(CoutP, nChunks);
  const int row_cols = fold ? 2 * CoutP : CoutP;   // TMEM columns per output row
  // accumulator ring: as many sets as fit the 512 TMEM columns (2 at N = 128, 8 for the narrow layers), or the tuning cap
  int nacc = 512 / (R * row_cols);
  nacc = nacc < 1 ? 1 : (nacc > MAX_ACC ? MAX_ACC : nacc);
  if (tuning().conv_nacc > 0 && nacc > tuning().conv_nacc) nacc = tuning().conv_nacc;
  int cols = 32;
  while (cols < nacc * R * row_cols) cols *= 2;
  const int tilesX = (OW + MT - 1) / MT, tilesY = (OH + R - 1) / R;
  // split-K when the caller lent a workspace (plain grids only): one launch covers the whole tiles (normal epilogue) and the
  // parts of the split tiles (raw sums into the workspace), a second one reduces the split region
  SplitK sk = {1, 0, 0, 0, 0, 0, nullptr, 0u, 0u, 3};
  if (ws != nullptr && ext == 0) {
    sk = plan_split(N, Cin, H, W, Cout, stride, dil, tuning().conv_grid_cap);
    if (sk.k > 1 && ws_bytes < sk.k * sk.part_stride * 4) sk.k = 1;
    sk.ws = ws;
  }
  const long long tiles = (long long)N * tilesX * tilesY;
  sk.as_wide = as_wide;
  if (sk.k <= 1) sk.from = (int)tiles;
  {
    const unsigned long long tp = (unsigned long long)tilesX * tilesY;
    if ((unsigned long long)tiles * tp < (1ull << 32) && tp * tilesX < (1ull << 32)) {   // q = umulhi(t, ceil(2^32 / d)) exact for t d < 2^32
      sk.magic_tp = (uint32_t)(((1ull << 32) + tp - 1) / tp);
      sk.magic_tx = tilesX == 1 ? 0u : (uint32_t)(((1ull << 32) + tilesX - 1) / tilesX);
    }
    if (tp == 1) sk.magic_tp = 0;          // ceil(2^32 / 1) does not fit 32 bits: the kernel divides
  }
  const long long numTiles = sk.from + (tiles - sk.from) * sk.k;   // work items
  const int cap = tuning().conv_grid_cap > 0 ? tuning().conv_grid_cap : kNumSMs;
  const unsigned grid = (unsigned)(numTiles < cap ? numTiles : cap);
#define MFN_UMMA_LAUNCH(FOLD_, TPS_)                                                                                       \
  conv3x3_umma_kernel<FOLD_, TPS_><<<grid, NTHREADS, sm.total, st>>>(x, x_bs, wpack, bias, out, out_bs, Cin, H, W, OH, OW, Cout,  \
                                                                     CoutP, nChunks, slope, tilesX, tilesY, (int)numTiles,      \
                                                                     stride, dil, out_mode, cols, nacc, ext, sk)
  if (taps_per_stage(CoutP) == 9) MFN_UMMA_LAUNCH(true, 9);
  else if (taps_per_stage(CoutP) == 3) MFN_UMMA_LAUNCH(true, 3);
  else if (fold) MFN_UMMA_LAUNCH(true, 1);
  else MFN_UMMA_LAUNCH(false, 1);
#undef MFN_UMMA_LAUNCH
  const int rc = check_launch("conv3x3_umma_kernel");
  if (rc != 0 || sk.k <= 1) return rc;
  const long long total = sk.part_stride;
  long long blocks = (total + 255) / 256;
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  conv3x3_umma_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(sk, N - sk.n_lo, bias, out, out_bs, Cout, OH, OW, slope, out_mode);
  return check_launch("conv3x3_umma_reduce_kernel");
}

}  // namespace mfn
