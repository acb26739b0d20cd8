// warp_mma.cu -- fused flow-guided feature warp (K3) on tensor cores, fp32-accurate (bf16 hi/lo split, 3 MMAs per product).
//
// Same math as deform_fwd_kernel<.., SHARED=true> in warp_fwd.cu (network/MaskFlownet.py:228-233 / layer.py:117-124):
//   flow = Upsample(up)(flow_c); mask = Upsample(up)(mask_c)
//   conv[f] = b[f] + sum_{c,tap} W[f,c,tap] * bilinear(x[c], y - 1 + ky + dy, x - 1 + kx + dx),  (dy,dx) = flow*scale/stride
//   out = LeakyReLU(conv * sigmoid(mask) + tradeoff)
// but as an implicit GEMM (M = pixels, N = F, K = 9*C) on the pipeline of conv3x3.cu's per-tap-tile variant: for every
// (32-channel chunk, tap) the CTA builds the A tile [rows x 32 px][32 ch] by bilinear GATHER (lane = pixel; its four corner
// offsets / weights for the tap are computed once and reused for all channels), splits it to bf16 hi/lo in swizzled shared
// memory, and multiplies it with the pre-packed weight tile of that (chunk, tap).  The im2col never exists in HBM.
#include "mma_tiles.cuh"

namespace mfn {

struct AxisW {
  int i0, i1;
  float w0, w1;
};
template <int BORDER>
__device__ __forceinline__ AxisW axis_w(float c, int n) {
  AxisW a;
  if (BORDER == MFN_BORDER_MXNET15) {
    const bool valid = (c >= 0.f) && (c < (float)n);
    int c0 = (int)floorf(c);
    float l;
    if (c0 >= n - 1) {
      c0 = n - 1;
      a.i1 = c0;
      l = 0.f;
    } else {
      a.i1 = c0 + 1;
      l = c - (float)c0;
    }
    a.i0 = c0;
    a.w0 = 1.f - l;
    a.w1 = l;
    if (!valid) {
      a.i0 = a.i1 = 0;
      a.w0 = a.w1 = 0.f;
    }
  } else {
    const bool valid = (c > -1.f) && (c < (float)n);
    const int c0 = (int)floorf(c);
    const float l = c - (float)c0;
    a.w0 = (valid && c0 >= 0) ? 1.f - l : 0.f;
    a.w1 = (valid && c0 + 1 <= n - 1) ? l : 0.f;
    a.i0 = max(min(c0, n - 1), 0);
    a.i1 = max(min(c0 + 1, n - 1), 0);
  }
  return a;
}

// WC warp columns x NTN n8 tiles per warp cover F (padded); WR = 8 / WC image rows x 32 pixels per CTA.
template <int WC, int NTN, int BORDER>
__global__ void __launch_bounds__(c3::NTHREADS, 1)
    warp_mma_kernel(const float* __restrict__ x, const float* __restrict__ flow_c, const float* __restrict__ mask_c,
                    const unsigned char* __restrict__ wpack, const float* __restrict__ bias,
                    const float* __restrict__ tradeoff, float* __restrict__ out, float* __restrict__ flow_up_out,
                    float* __restrict__ mask_up_out, float* __restrict__ conv_out, int C, int H, int W, int F, int FP,
                    int nChunks, int up, float flow_scale, float level_stride, float slope, int tilesX, int tilesY) {
  using namespace c3;
  constexpr int WR = 8 / WC;
  constexpr int IN_LO = WR * TW * PXB;             // lo image offset inside an input stage
  constexpr int IN_STAGE = 2 * IN_LO;
  constexpr int UPW = WR * 4 / 8;                  // (row, 8-channel chunk) units per warp: 4 (WC=1) or 2 (WC=2)
  const int WT_LO = FP * PXB, WT_BYTES = 2 * WT_LO;

  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* in_s = smem;                       // [2][IN_STAGE]
  unsigned char* wt_s = smem + 2 * IN_STAGE;        // [WSTAGES][WT_BYTES]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wr = warp / WC, wc = warp % WC;
  const int tile = blockIdx.x;
  const int tx = tile % tilesX, ty = (tile / tilesX) % tilesY, n = tile / (tilesX * tilesY);
  const int x0 = tx * TW, y0 = ty * WR;
  const size_t plane = (size_t)H * W;
  const float* xn = x + (size_t)n * C * plane;

  // ---- gather role: this warp's units all belong to ONE tile row: lane = pixel (grow, x0 + lane) ----
  const int grow = (warp * UPW) >> 2;               // tile row of the gathered pixel
  const int gch0 = (warp * UPW) & 3;                // first 8-channel chunk handled by this warp
  const int gy = y0 + grow, gx = x0 + lane;
  const bool gin = gy < H && gx < W;
  float dy = 0.f, dx = 0.f;
  if (gin) {
    const int Hc = H / up, Wc = W / up;
    const float* fc = flow_c + (size_t)n * 2 * Hc * Wc;
    const float fy = upsample_at(fc, Hc, Wc, up, gy, gx);
    const float fx = upsample_at(fc + (size_t)Hc * Wc, Hc, Wc, up, gy, gx);
    // offsets exactly as the reference rounds them: (flow * scale) / stride   (MaskFlownet.py:230)
    dy = __fdiv_rn(__fmul_rn(fy, flow_scale), level_stride);
    dx = __fdiv_rn(__fmul_rn(fx, flow_scale), level_stride);
    if (gch0 == 0) {   // one warp per tile row publishes the up-sampled flow / mask
      if (flow_up_out) {
        flow_up_out[((size_t)n * 2) * plane + (size_t)gy * W + gx] = fy;
        flow_up_out[((size_t)n * 2 + 1) * plane + (size_t)gy * W + gx] = fx;
      }
      if (mask_up_out && mask_c)
        mask_up_out[(size_t)n * plane + (size_t)gy * W + gx] = upsample_at(mask_c + (size_t)n * Hc * Wc, Hc, Wc, up, gy, gx);
    }
  }
  auto gather_tile = [&](int it, float (&e)[UPW][8]) {
    const int q = it / 9, tap = it - 9 * q;
    const int ky = tap / 3, kx = tap - 3 * ky;
    const AxisW hA = axis_w<BORDER>((float)(gy - 1 + ky) + dy, H);
    const AxisW wA = axis_w<BORDER>((float)(gx - 1 + kx) + dx, W);
    const float w00 = hA.w0 * wA.w0, w01 = hA.w0 * wA.w1, w10 = hA.w1 * wA.w0, w11 = hA.w1 * wA.w1;
    const int o00 = hA.i0 * W + wA.i0, o01 = hA.i0 * W + wA.i1, o10 = hA.i1 * W + wA.i0, o11 = hA.i1 * W + wA.i1;
    const bool any = gin && ((w00 != 0.f) || (w01 != 0.f) || (w10 != 0.f) || (w11 != 0.f));
#pragma unroll
    for (int k = 0; k < UPW; ++k) {
      const int c0 = 32 * q + 8 * (gch0 + k);
      const float* pl = xn + (size_t)c0 * plane;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float s = 0.f;
        if (any && c0 + c < C)   // same association as the oracle: ((w00*v00 + w01*v01) + w10*v10) + w11*v11
          s = w00 * __ldg(pl + o00) + w01 * __ldg(pl + o01) + w10 * __ldg(pl + o10) + w11 * __ldg(pl + o11);
        e[k][c] = s;
        pl += plane;
      }
    }
  };
  auto store_tile = [&](int stage, const float (&e)[UPW][8]) {
#pragma unroll
    for (int k = 0; k < UPW; ++k) {
      uint4 hi, lo;
      split_pair(e[k][0], e[k][1], hi.x, lo.x);
      split_pair(e[k][2], e[k][3], hi.y, lo.y);
      split_pair(e[k][4], e[k][5], hi.z, lo.z);
      split_pair(e[k][6], e[k][7], hi.w, lo.w);
      unsigned char* dst = in_s + stage * IN_STAGE + grow * (TW * PXB) + swz(lane, gch0 + k);
      *reinterpret_cast<uint4*>(dst) = hi;
      *reinterpret_cast<uint4*>(dst + IN_LO) = lo;
    }
  };
  auto load_weights = [&](int it, int stage) {
    const unsigned char* src = wpack + (size_t)it * WT_BYTES;
    const uint32_t dst = smem_u32(wt_s + stage * WT_BYTES);
    for (int o = tid * 16; o < WT_BYTES; o += NTHREADS * 16) cp_async16(dst + o, src + o);
  };

  // ---- MMA role ----
  const int g = lane >> 2, j = lane & 3;
  const int l8 = lane & 7, mi = lane >> 3;
  const int swl = (l8 >> 1) & 3;
  uint32_t offA[2][2], offB[2];
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
      offA[mt][kk] = (uint32_t)((16 * mt + 8 * (mi & 1) + l8) * PXB + (((2 * kk + (mi >> 1)) ^ swl) << 4));
    offB[kk] = (uint32_t)((8 * (mi >> 1) + l8) * PXB + (((2 * kk + (mi & 1)) ^ swl) << 4));
  }
  const int fbase = wc * NTN * 8;

  float acc[2][NTN][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTN; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

  const int nIter = nChunks * 9;
  {
    float e[UPW][8];
    gather_tile(0, e);
    store_tile(0, e);
  }
  load_weights(0, 0);
  cp_async_commit();
  if (nIter > 1) load_weights(1, 1);
  cp_async_commit();

  const uint32_t in_u32 = smem_u32(in_s), wt_u32 = smem_u32(wt_s);
  float pe[UPW][8];
  for (int it = 0; it < nIter; ++it) {
    const int q = it / 9;
    cp_async_wait<1>();
    __syncthreads();
    if (it + 2 < nIter) load_weights(it + 2, (it + 2) % WSTAGES);
    cp_async_commit();
    if (it + 1 < nIter) gather_tile(it + 1, pe);   // next tap's samples: in flight during this tap's MMAs

    const uint32_t wst = wt_u32 + (uint32_t)((it % WSTAGES) * WT_BYTES);
    const uint32_t ist = in_u32 + (uint32_t)((it & 1) * IN_STAGE + wr * (TW * PXB));
    const bool half_chunk = 32 * q + 16 >= C;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      if (kk == 1 && half_chunk) break;
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        ldsm_x4(ist + offA[mt][kk], ah[mt]);
        ldsm_x4(ist + offA[mt][kk] + IN_LO, al[mt]);
      }
#pragma unroll
      for (int nt = 0; nt < NTN; nt += 2) {
        uint32_t bh[4], bl[4];
        const uint32_t b = wst + (uint32_t)((fbase + 8 * nt) * PXB) + offB[kk];
        ldsm_x4(b, bh);
        ldsm_x4(b + WT_LO, bl);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_bf16(acc[mt][nt], ah[mt], bl[0], bl[1]);
          mma_bf16(acc[mt][nt], al[mt], bh[0], bh[1]);
          mma_bf16(acc[mt][nt], ah[mt], bh[0], bh[1]);
          if (nt + 1 < NTN) {
            mma_bf16(acc[mt][nt + 1], ah[mt], bl[2], bl[3]);
            mma_bf16(acc[mt][nt + 1], al[mt], bh[2], bh[3]);
            mma_bf16(acc[mt][nt + 1], ah[mt], bh[2], bh[3]);
          }
        }
      }
    }
    if (it + 1 < nIter) store_tile((it + 1) & 1, pe);
  }

  // ---- epilogue: (conv + bias) * sigmoid(mask) + tradeoff -> LeakyReLU ----
  const int y = y0 + wr;
  if (y >= H) return;
  float sig[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int xx = x0 + 16 * mt + g + 8 * h;
      sig[mt][h] = 1.f;
      if (mask_c && xx < W) {
        const int Hc = H / up, Wc = W / up;
        sig[mt][h] = sigmoidf_(upsample_at(mask_c + (size_t)n * Hc * Wc, Hc, Wc, up, y, xx));
      }
    }
  const size_t rowbase = (size_t)n * F * plane + (size_t)y * W;
#pragma unroll
  for (int nt = 0; nt < NTN; ++nt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int f = fbase + 8 * nt + 2 * j + (i & 1);
      if (f >= F) continue;
      const float b = bias ? __ldg(bias + f) : 0.f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int xx = x0 + 16 * mt + g + 8 * (i >> 1);
        if (xx >= W) continue;
        const size_t oi = rowbase + (size_t)f * plane + xx;
        float v = acc[mt][nt][i] + b;
        if (conv_out) conv_out[oi] = v;
        v *= sig[mt][i >> 1];
        if (tradeoff) v += __ldg(tradeoff + oi);
        out[oi] = leaky(v, slope);
      }
    }
  }
}

template <int WC, int NTN, int BORDER>
static int launch_warp_mma(const float* x, const float* flow_c, const float* mask_c, const unsigned char* wpack,
                           const float* bias, const float* tradeoff, float* out, float* fup, float* mup, float* conv_out,
                           int N, int C, int H, int W, int F, int up, float fs, float ls, float slope, cudaStream_t st) {
  using namespace c3;
  constexpr int WR = 8 / WC;
  const int FP = cout_pad(F), nChunks = (C + 31) / 32;
  const int tilesX = (W + TW - 1) / TW, tilesY = (H + WR - 1) / WR;
  const int smem = 2 * (2 * WR * TW * PXB) + WSTAGES * 2 * FP * PXB;
  static SmemOptIn opt;
  {
    const cudaError_t e = ensure_dyn_smem(warp_mma_kernel<WC, NTN, BORDER>, smem, opt);
    if (e != cudaSuccess) return fail((int)e, "cudaFuncSetAttribute(warp_mma_kernel): %s", cudaGetErrorString(e));
  }
  const unsigned grid = (unsigned)((long long)N * tilesX * tilesY);
  warp_mma_kernel<WC, NTN, BORDER><<<grid, NTHREADS, smem, st>>>(x, flow_c, mask_c, wpack, bias, tradeoff, out, fup, mup,
                                                                conv_out, C, H, W, F, FP, nChunks, up, fs, ls, slope,
                                                                tilesX, tilesY);
  return check_launch("warp_mma_kernel");
}

template <int BORDER>
static int dispatch_warp_mma(const float* x, const float* flow_c, const float* mask_c, const unsigned char* wpack,
                             const float* bias, const float* tradeoff, float* out, float* fup, float* mup,
                             float* conv_out, int N, int C, int H, int W, int F, int up, float fs, float ls, float slope,
                             cudaStream_t st) {
  const int nt = (F + 7) / 8;
#define MFN_WM(WC_, NTN_)                                                                                             \
  launch_warp_mma<WC_, NTN_, BORDER>(x, flow_c, mask_c, wpack, bias, tradeoff, out, fup, mup, conv_out, N, C, H, W, F, up, \
                                     fs, ls, slope, st)
  if (nt <= 4) return MFN_WM(1, 4);
  if (nt <= 8) return MFN_WM(1, 8);
  if (nt <= 12) return MFN_WM(2, 6);
  return MFN_WM(2, 8);
#undef MFN_WM
}

}  // namespace mfn

extern "C" int mfn_warp_mask_forward_tc(const float* x, const float* flow_coarse, const float* mask_coarse,
                                        const void* packed_weight, const float* bias, const float* tradeoff, float* out,
                                        float* flow_up_out, float* mask_up_out, float* conv_out, int N, int C, int H, int W,
                                        int F, int upsample_factor, float flow_scale, float level_stride,
                                        float leaky_slope, int border_mode, void* stream) {
  using namespace mfn;
  MFN_REQUIRE(x && flow_coarse && packed_weight && out, MFN_ERR_INVALID_ARG, "mfn_warp_mask_forward_tc: null pointer");
  MFN_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && F > 0, MFN_ERR_INVALID_ARG,
              "mfn_warp_mask_forward_tc: non-positive extent");
  MFN_REQUIRE(F <= 128, MFN_ERR_UNSUPPORTED, "mfn_warp_mask_forward_tc: at most 128 output channels (got %d)", F);
  MFN_REQUIRE(upsample_factor >= 1 && H % upsample_factor == 0 && W % upsample_factor == 0, MFN_ERR_INVALID_ARG,
              "mfn_warp_mask_forward_tc: H and W must be multiples of upsample_factor");
  MFN_REQUIRE(level_stride > 0.f, MFN_ERR_INVALID_ARG, "mfn_warp_mask_forward_tc: level_stride must be positive");
  MFN_REQUIRE(border_mode == MFN_BORDER_MXNET15 || border_mode == MFN_BORDER_ZERO_CORNER, MFN_ERR_INVALID_ARG,
              "mfn_warp_mask_forward_tc: unknown border_mode %d", border_mode);
  MFN_REQUIRE(aligned(packed_weight, 16), MFN_ERR_ALIGNMENT, "mfn_warp_mask_forward_tc: packed weights must be 16-byte aligned");
  MFN_REQUIRE((long long)C * H * W < (1LL << 31) && (long long)F * H * W < (1LL << 31), MFN_ERR_ALIGNMENT,
              "mfn_warp_mask_forward_tc: extents overflow kernel indexing");
  const unsigned char* wp = static_cast<const unsigned char*>(packed_weight);
  cudaStream_t st = as_stream(stream);
  if (border_mode == MFN_BORDER_MXNET15)
    return dispatch_warp_mma<MFN_BORDER_MXNET15>(x, flow_coarse, mask_coarse, wp, bias, tradeoff, out, flow_up_out,
                                                 mask_up_out, conv_out, N, C, H, W, F, upsample_factor, flow_scale,
                                                 level_stride, leaky_slope, st);
  return dispatch_warp_mma<MFN_BORDER_ZERO_CORNER>(x, flow_coarse, mask_coarse, wp, bias, tradeoff, out, flow_up_out,
                                                   mask_up_out, conv_out, N, C, H, W, F, upsample_factor, flow_scale,
                                                   level_stride, leaky_slope, st);
}
