// warp_lin.cu -- K3 (flow-guided deformable warp x occlusion mask) evaluated EXACTLY through linearity, for every pixel and
// both border rules; no per-pixel fallback list, cost independent of the flow field.
//
// Reference semantics: network/MaskFlownet.py:228-233 -> network/layer.py:117-124 (DeformableConvolution, 3x3, all nine tap
// offsets equal to the up-sampled flow), then * sigmoid(mask) + trade-off -> LeakyReLU.
//
// Every tap samples x at (h_i, w_j) = (y - 1 + i + dy, x - 1 + j + dx) with a SEPARABLE rule: the sample is a(h_i)^T X b(w_j),
// where the per-axis coefficient vector is
//   zero-corner rule (DCNv2 / torchvision, MFN_BORDER_ZERO_CORNER):  a = a_Z = the hat function on the zero-extended axis;
//   MXNet-1.5 rule (MFN_BORDER_MXNET15):  a = a_Z + a_D with a_D(h) = -(1 + h) e_0       for h in (-1, 0)   (sample forced to 0)
//                                                            a_D(h) = (h - (n-1)) e_{n-1} for h in (n-1, n) (collapsed on the last pixel)
//                                                            a_D(h) = 0 elsewhere.
// Summing over the taps with the weights W_ij:
//   Z x Z : bilinear sample at (y + dy, x + dx) of Yext = conv3x3(x, W) evaluated on the grid extended by one pixel per side
//           (tcgen05 kernel, conv3x3_umma.cu with ext = 1) -- all of the operator in zero-corner mode;
//   D x Z : sum_i alpha_i * lerp_w( Rrow[B_i][i] )   with Rrow[B][i] = conv1d(x[row B], W[i, :]) on the extended row,
//   Z x D : sum_j beta_j  * lerp_h( Rcol[B_j][j] )   with Rcol[B][j] = conv1d(x[:, col B], W[:, j]),
//   D x D : sum_ij alpha_i beta_j T[B_i][B_j][i][j]  with T = W_ij . x[:, row B_i, col B_j]
// (B = first / last row or column).  The three tables are tiny (one launch); only pixels whose taps reach the one-pixel bands
// read them.  Band membership uses the same float expression (y - 1 + i) + dy as the reference kernel, so the operator's
// discontinuities (h = 0, h = H) are taken on the same side as the tap-by-tap kernels and the oracle.
#include "mma_tiles.cuh"

namespace mfn {
namespace wl {

struct Band {
  float a[3];   // correction coefficient of tap row / column i (0 = not in a band)
  int B[3];     // 0 = first row / column, 1 = last
  bool any;
};
__device__ __forceinline__ Band bands(int p, float d, int n) {
  Band b;
  b.any = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float h = (float)(p - 1 + i) + d;
    float a = 0.f;
    int B = 0;
    if (h > -1.f && h < 0.f) {
      a = -(1.f + h);
    } else if (h > (float)(n - 1) && h < (float)n) {
      a = h - (float)(n - 1);
      B = 1;
    }
    b.a[i] = a;
    b.B[i] = B;
    b.any = b.any || (a != 0.f);
  }
  return b;
}

// linear interpolation of a zero-extended table row t[0 .. n+1] (entry v <-> position v - 1) at position q
__device__ __forceinline__ float lerp_ext(const float* __restrict__ t, int n, float q) {
  const float fl = floorf(q);
  const int i0 = (int)fl + 1;
  const float l = q - fl;
  float v = 0.f;
  if (i0 >= 0 && i0 <= n + 1) v += (1.f - l) * __ldg(t + i0);
  if (i0 + 1 >= 0 && i0 + 1 <= n + 1) v += l * __ldg(t + i0 + 1);
  return v;
}

// Rrow[n][B][i][f][W+2], Rcol[n][B][j][f][H+2], T[n][Br][Bc][i][j][f]
__global__ void __launch_bounds__(256)
    warp_tables_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ Rrow,
                       float* __restrict__ Rcol, float* __restrict__ T, int N, int C, int H, int W, int F) {
  const long long nRow = (long long)N * 6 * F * (W + 2), nCol = (long long)N * 6 * F * (H + 2), nT = (long long)N * 36 * F;
  const size_t plane = (size_t)H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < nRow + nCol + nT;
       idx += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (idx < nRow) {
      const int v = (int)(idx % (W + 2));
      const int f = (int)((idx / (W + 2)) % F);
      const int i = (int)((idx / ((long long)(W + 2) * F)) % 3);
      const int B = (int)((idx / ((long long)(W + 2) * F * 3)) % 2);
      const int n = (int)(idx / ((long long)(W + 2) * F * 6));
      const float* xr = x + (size_t)n * C * plane + (size_t)(B ? H - 1 : 0) * W;
      const float* wf = w + (size_t)f * C * 9 + 3 * i;
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int xx = v - 1 + j - 1;
          if (xx >= 0 && xx < W) acc = fmaf(__ldg(wf + (size_t)c * 9 + j), __ldg(xr + (size_t)c * plane + xx), acc);
        }
      }
      Rrow[idx] = acc;
    } else if (idx < nRow + nCol) {
      const long long k = idx - nRow;
      const int u = (int)(k % (H + 2));
      const int f = (int)((k / (H + 2)) % F);
      const int j = (int)((k / ((long long)(H + 2) * F)) % 3);
      const int B = (int)((k / ((long long)(H + 2) * F * 3)) % 2);
      const int n = (int)(k / ((long long)(H + 2) * F * 6));
      const float* xc = x + (size_t)n * C * plane + (B ? W - 1 : 0);
      const float* wf = w + (size_t)f * C * 9 + j;
      for (int c = 0; c < C; ++c) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int yy = u - 1 + i - 1;
          if (yy >= 0 && yy < H) acc = fmaf(__ldg(wf + (size_t)c * 9 + 3 * i), __ldg(xc + (size_t)c * plane + (size_t)yy * W), acc);
        }
      }
      Rcol[k] = acc;
    } else {
      const long long k = idx - nRow - nCol;
      const int f = (int)(k % F);
      const int ij = (int)((k / F) % 9);
      const int Bc = (int)((k / ((long long)F * 9)) % 2);
      const int Br = (int)((k / ((long long)F * 18)) % 2);
      const int n = (int)(k / ((long long)F * 36));
      const float* xp = x + (size_t)n * C * plane + (size_t)(Br ? H - 1 : 0) * W + (Bc ? W - 1 : 0);
      const float* wf = w + (size_t)f * C * 9 + ij;
      for (int c = 0; c < C; ++c) acc = fmaf(__ldg(wf + (size_t)c * 9), __ldg(xp + (size_t)c * plane), acc);
      T[k] = acc;
    }
  }
}

// One thread per pixel, loop over the F output channels (all accesses coalesced along x).
template <int BORDER>
__global__ void __launch_bounds__(256)
    warp_lin_kernel(const float* __restrict__ Yext, const float* __restrict__ Rrow, const float* __restrict__ Rcol,
                    const float* __restrict__ T, const float* __restrict__ flow_c, const float* __restrict__ mask_c,
                    const float* __restrict__ bias, const float* __restrict__ tradeoff, float* __restrict__ out,
                    float* __restrict__ flow_up_out, float* __restrict__ mask_up_out, int N, int H, int W, int F, int up,
                    float flow_scale, float level_stride, float slope) {
  const long long total = (long long)N * H * W;
  const size_t plane = (size_t)H * W;
  const int WE = W + 2, HE = H + 2;
  const size_t eplane = (size_t)HE * WE;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int xq = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / plane);
    const int Hc = H / up, Wc = W / up;
    const float* fc = flow_c + (size_t)n * 2 * Hc * Wc;
    const float fy = upsample_at(fc, Hc, Wc, up, y, xq);
    const float fx = upsample_at(fc + (size_t)Hc * Wc, Hc, Wc, up, y, xq);
    const float mask_v = mask_c ? upsample_at(mask_c + (size_t)n * Hc * Wc, Hc, Wc, up, y, xq) : 0.f;
    const size_t pix = (size_t)y * W + xq;
    if (flow_up_out) {
      flow_up_out[((size_t)n * 2 + 0) * plane + pix] = fy;
      flow_up_out[((size_t)n * 2 + 1) * plane + pix] = fx;
    }
    if (mask_up_out && mask_c) mask_up_out[(size_t)n * plane + pix] = mask_v;
    // offsets exactly as the reference rounds them: (flow * scale) / stride   (MaskFlownet.py:230)
    const float dy = __fdiv_rn(__fmul_rn(fy, flow_scale), level_stride);
    const float dx = __fdiv_rn(__fmul_rn(fx, flow_scale), level_stride);
    const float h0 = (float)y + dy, w0 = (float)xq + dx;
    // ---- Z x Z: bilinear sample of the extended convolution (entry (u, v) <-> position (u - 1, v - 1)) ----
    const float fh = floorf(h0), fw = floorf(w0);
    // far outside: keep the integer conversion defined
    const bool farout = !(h0 > -3.f && h0 < (float)(H + 2) && w0 > -3.f && w0 < (float)(W + 2));
    const int i0 = farout ? -8 : (int)fh + 1, j0 = farout ? -8 : (int)fw + 1;
    const float lh = h0 - fh, lw = w0 - fw;
    const bool r0 = i0 >= 0 && i0 <= H + 1, r1 = i0 + 1 >= 0 && i0 + 1 <= H + 1;
    const bool c0 = j0 >= 0 && j0 <= W + 1, c1 = j0 + 1 >= 0 && j0 + 1 <= W + 1;
    const float w00 = (r0 && c0) ? (1.f - lh) * (1.f - lw) : 0.f, w01 = (r0 && c1) ? (1.f - lh) * lw : 0.f;
    const float w10 = (r1 && c0) ? lh * (1.f - lw) : 0.f, w11 = (r1 && c1) ? lh * lw : 0.f;
    const int ic0 = min(max(i0, 0), H + 1), ic1 = min(max(i0 + 1, 0), H + 1);
    const int jc0 = min(max(j0, 0), W + 1), jc1 = min(max(j0 + 1, 0), W + 1);
    const int o00 = ic0 * WE + jc0, o01 = ic0 * WE + jc1, o10 = ic1 * WE + jc0, o11 = ic1 * WE + jc1;
    const bool anyz = (w00 != 0.f) || (w01 != 0.f) || (w10 != 0.f) || (w11 != 0.f);
    Band bh, bw;
    bh.any = bw.any = false;
    if (BORDER == MFN_BORDER_MXNET15) {
      bh = bands(y, dy, H);
      bw = bands(xq, dx, W);
    }
    const float sig = mask_c ? sigmoidf_(mask_v) : 1.f;
    const float* yp = Yext + (size_t)n * F * eplane;
    float* op = out + (size_t)n * F * plane + pix;
    const float* tp = tradeoff ? tradeoff + (size_t)n * F * plane + pix : nullptr;
    const float* rrow = Rrow + (size_t)n * 6 * F * WE;
    const float* rcol = Rcol + (size_t)n * 6 * F * HE;
    const float* tt = T + (size_t)n * 36 * F;
#pragma unroll 4
    for (int f = 0; f < F; ++f) {
      const float* q = yp + (size_t)f * eplane;
      float v = 0.f;
      if (anyz) v = w00 * __ldg(q + o00) + w01 * __ldg(q + o01) + w10 * __ldg(q + o10) + w11 * __ldg(q + o11);
      if (BORDER == MFN_BORDER_MXNET15 && (bh.any || bw.any)) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (bh.a[i] != 0.f) v += bh.a[i] * lerp_ext(rrow + ((size_t)(bh.B[i] * 3 + i) * F + f) * WE, W, w0);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (bw.a[j] != 0.f) v += bw.a[j] * lerp_ext(rcol + ((size_t)(bw.B[j] * 3 + j) * F + f) * HE, H, h0);
        if (bh.any && bw.any) {
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (bh.a[i] != 0.f && bw.a[j] != 0.f)
                v += bh.a[i] * bw.a[j] * __ldg(tt + ((size_t)((bh.B[i] * 2 + bw.B[j]) * 9 + 3 * i + j)) * F + f);
        }
      }
      if (bias) v += __ldg(bias + f);
      v *= sig;
      if (tp) v += __ldg(tp + (size_t)f * plane);
      op[(size_t)f * plane] = leaky(v, slope);
    }
  }
}

}  // namespace wl

long long warp_lin_workspace_bytes(int N, int F, int H, int W) {
  const long long ye = (long long)N * F * (H + 2) * (W + 2), tr = (long long)N * 6 * F * (W + 2), tc = (long long)N * 6 * F * (H + 2),
                  t = (long long)N * 36 * F;
  return (ye + tr + tc + t) * 4 + 64;
}

// returns -1 when the extended tcgen05 convolution does not fit the shape (caller uses the list-based path)
int launch_warp_lin(const float* x, const float* flow_c, const float* mask_c, const float* weight, const void* packed_weight,
                    const float* bias, const float* tradeoff, void* workspace, float* out, float* fup, float* mup, int N,
                    int C, int H, int W, int F, int up, float fs, float ls, float slope, int border_mode, cudaStream_t st) {
  using namespace wl;
  float* Yext = static_cast<float*>(workspace);
  float* Rrow = Yext + (size_t)N * F * (H + 2) * (W + 2);
  float* Rcol = Rrow + (size_t)N * 6 * F * (W + 2);
  float* T = Rcol + (size_t)N * 6 * F * (H + 2);
  const unsigned char* wp = static_cast<const unsigned char*>(packed_weight);
  int rc = conv3x3_umma_launch(x, (long long)C * H * W, wp + conv3x3_sync_packed_bytes(C, F), nullptr, Yext,
                               (long long)F * (H + 2) * (W + 2), N, C, H, W, F, 1, 1, MFN_CONV_OUT_NCHW, 1.0f, st, 1);
  if (rc) return rc;
  if (border_mode == MFN_BORDER_MXNET15) {
    const long long tot = (long long)N * 6 * F * (W + 2) + (long long)N * 6 * F * (H + 2) + (long long)N * 36 * F;
    long long blocks = (tot + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    warp_tables_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, weight, Rrow, Rcol, T, N, C, H, W, F);
    rc = check_launch("warp_tables_kernel");
    if (rc) return rc;
  }
  const long long total = (long long)N * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (border_mode == MFN_BORDER_MXNET15)
    warp_lin_kernel<MFN_BORDER_MXNET15><<<(unsigned)blocks, 256, 0, st>>>(Yext, Rrow, Rcol, T, flow_c, mask_c, bias, tradeoff, out,
                                                                         fup, mup, N, H, W, F, up, fs, ls, slope);
  else
    warp_lin_kernel<MFN_BORDER_ZERO_CORNER><<<(unsigned)blocks, 256, 0, st>>>(Yext, Rrow, Rcol, T, flow_c, mask_c, bias, tradeoff,
                                                                             out, fup, mup, N, H, W, F, up, fs, ls, slope);
  return check_launch("warp_lin_kernel");
}

}  // namespace mfn
