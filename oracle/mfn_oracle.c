/*
 * mfn_oracle.c -- CPU oracle for the MaskFlownet hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file restates, in plain C, the semantics of the Apache MXNet 1.5 operators that the
 * reference calls on its hot path.  It is the *checker* for the CUDA kernels in
 * maskflownet_b200/csrc; nothing in the product path may link, import or call it (only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do).
 *
 * The arithmetic of this path is NOT in /root/reference: it lives in the un-vendored third-party
 * dependency `mxnet` ("MXNet 1.5", reference README.md:27).  Each function below cites the reference
 * call site whose parameters it follows and restates the published MXNet operator definition:
 *
 *   mfn_ref_correlation_forward/backward   F.Correlation            network/MaskFlownet.py:193-195, 440-441
 *   mfn_ref_deformable_conv_forward        F.contrib.DeformableConvolution   network/layer.py:117-124
 *   mfn_ref_upsample                       Upsample(f) block        network/MaskFlownet.py:35-62
 *   mfn_ref_grid_generator_warp            F.GridGenerator('warp')  network/layer.py:17
 *   mfn_ref_bilinear_sampler               F.BilinearSampler        network/layer.py:18
 *
 * PARITY PIN STATUS: the reference ships no tests / golden vectors and MXNet cannot be installed
 * here, so this oracle is pinned against independent implementations available in this container
 * (TVM topi `correlation_nchw_python`, torchvision `deform_conv2d` in the interior and in
 * zero-corner mode, TVM topi `deformable_conv2d_nchw_python` for the lower half of the MXNet-1.5
 * border rule (taps in (-1,0) contribute zero), torch `grid_sample(align_corners=True)`, torch
 * `conv_transpose2d` for Upsample) -- see tests/golden/make_golden*.py and DESIGN.md.
 * Against MXNet itself: "parity unpinned".
 *
 * Build: see oracle/Makefile (gcc -O2 [-fopenmp]).  All tensors are fp32, NCHW, contiguous.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MFN_REF_API __attribute__((visibility("default")))

MFN_REF_API int mfn_ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

MFN_REF_API void mfn_ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Correlation (MXNet src/operator/correlation.cc, CorrelationForward -- restated from its definition).
 *
 * Steps of the MXNet CPU operator, kept literally because they are what "the reference CPU path"
 * executes (BASELINE.json configs[0]):
 *   1. both inputs are copied into zero-padded channel-last temporaries  tmp[n][y+pad][x+pad][c]
 *   2. for every output row i, column j, sample n and displacement channel q:
 *        (x1,y1) = (j*stride1 + md, i*stride1 + md)                    centre in padded data1
 *        (x2,y2) = (x1 + (q % G - r)*stride2, y1 + (q / G - r)*stride2) centre in padded data2
 *        out += sum_{h,w<k} sum_c  t1[y1+h][x1+w][c] (*|-) t2[y2+h][x2+w][c]
 *   3. out /= k*k*C
 * with r = md/stride2, G = 2r+1, border = md + (k-1)/2,
 *      top_h = ceil((H + 2 pad - 2 border)/stride1), same for w.
 * Reference call sites use pad=md, k=1, stride1=stride2=1, is_multiply=1 (MaskFlownet.py:195,441).
 * `threads` = 1 reproduces MXNet's single-threaded loop nest; >1 distributes output rows (n,i).
 * ------------------------------------------------------------------------------------------------ */
static void pad_to_channel_last(const float* src, float* dst, int N, int C, int H, int W, int pad) {
  const int PH = H + 2 * pad, PW = W + 2 * pad;
  memset(dst, 0, sizeof(float) * (size_t)N * PH * PW * C);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y) {
        const float* s = src + (((size_t)n * C + c) * H + y) * W;
        float* d = dst + (((size_t)n * PH + (y + pad)) * PW + pad) * C + c;
        for (int x = 0; x < W; ++x) d[(size_t)x * C] = s[x];
      }
}

MFN_REF_API int mfn_ref_correlation_out_shape(int H, int W, int pad_size, int kernel_size,
                                              int max_displacement, int stride1, int stride2,
                                              int* out_c, int* out_h, int* out_w) {
  if (kernel_size < 1 || (kernel_size & 1) == 0 || stride1 < 1 || stride2 < 1) return -1;
  const int kr = (kernel_size - 1) / 2;
  const int border = max_displacement + kr;
  const int ph = H + 2 * pad_size, pw = W + 2 * pad_size;
  const int th = (int)ceilf((float)(ph - 2 * border) / (float)stride1);
  const int tw = (int)ceilf((float)(pw - 2 * border) / (float)stride1);
  const int r = max_displacement / stride2;
  const int G = 2 * r + 1;
  if (th < 1 || tw < 1) return -1;
  *out_c = G * G;
  *out_h = th;
  *out_w = tw;
  return 0;
}

MFN_REF_API int mfn_ref_correlation_forward(const float* data1, const float* data2, float* out,
                                            int N, int C, int H, int W, int pad_size,
                                            int kernel_size, int max_displacement, int stride1,
                                            int stride2, int is_multiply, int threads) {
  int D, TH, TW;
  if (mfn_ref_correlation_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                    &D, &TH, &TW))
    return -1;
  const int PH = H + 2 * pad_size, PW = W + 2 * pad_size;
  const int r = max_displacement / stride2, G = 2 * r + 1;
  const size_t tmp_elems = (size_t)N * PH * PW * C;
  float* t1 = (float*)malloc(sizeof(float) * tmp_elems);
  float* t2 = (float*)malloc(sizeof(float) * tmp_elems);
  if (!t1 || !t2) {
    free(t1);
    free(t2);
    return -2;
  }
  pad_to_channel_last(data1, t1, N, C, H, W, pad_size);
  pad_to_channel_last(data2, t2, N, C, H, W, pad_size);
  const float norm = (float)(kernel_size * kernel_size * C);
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static) num_threads(threads > 0 ? threads : 1)
#endif
  for (int i = 0; i < TH; ++i)
    for (int n = 0; n < N; ++n)
      for (int j = 0; j < TW; ++j) {
        const int x1 = j * stride1 + max_displacement;
        const int y1 = i * stride1 + max_displacement;
        for (int q = 0; q < D; ++q) {
          const int x2 = x1 + (q % G - r) * stride2;
          const int y2 = y1 + (q / G - r) * stride2;
          float acc = 0.f;
          for (int h = 0; h < kernel_size; ++h)
            for (int w = 0; w < kernel_size; ++w) {
              /* positions outside the padded temporaries read as zero (cannot happen when
                 pad_size >= max_displacement + kernel_radius, the only regime the reference uses) */
              const int ya = y1 + h, xa = x1 + w, yb = y2 + h, xb = x2 + w;
              if (ya < 0 || ya >= PH || xa < 0 || xa >= PW) continue;
              const float* a = t1 + (((size_t)n * PH + ya) * PW + xa) * C;
              if (yb < 0 || yb >= PH || xb < 0 || xb >= PW) {
                if (!is_multiply)
                  for (int c = 0; c < C; ++c) acc += fabsf(a[c]);
                continue;
              }
              const float* b = t2 + (((size_t)n * PH + yb) * PW + xb) * C;
              if (is_multiply)
                for (int c = 0; c < C; ++c) acc += a[c] * b[c];
              else
                for (int c = 0; c < C; ++c) acc += fabsf(a[c] - b[c]);
            }
          out[(((size_t)n * D + q) * TH + i) * TW + j] = acc / norm;
        }
      }
  free(t1);
  free(t2);
  return 0;
}

/* Correlation backward for the regime the reference uses (k=1, strides 1, multiply, pad==md):
 *   g1[n,c,y,x] = 1/C * sum_q go[n,q,y,x]       * f2[n,c,y+dy,x+dx]     (zero outside)
 *   g2[n,c,y,x] = 1/C * sum_q go[n,q,y-dy,x-dx] * f1[n,c,y-dy,x-dx]     (terms outside dropped)
 * (MXNet CorrelationBackward restated).  Accumulates in double. */
MFN_REF_API int mfn_ref_correlation_backward(const float* grad_out, const float* data1,
                                             const float* data2, float* grad1, float* grad2, int N,
                                             int C, int H, int W, int max_displacement,
                                             int threads) {
  const int md = max_displacement, G = 2 * md + 1, D = G * G;
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static) num_threads(threads > 0 ? threads : 1)
#endif
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < C; ++c) {
      const float* f1 = data1 + ((size_t)n * C + c) * H * W;
      const float* f2 = data2 + ((size_t)n * C + c) * H * W;
      const float* go = grad_out + (size_t)n * D * H * W;
      float* g1 = grad1 + ((size_t)n * C + c) * H * W;
      float* g2 = grad2 + ((size_t)n * C + c) * H * W;
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          double a1 = 0.0, a2 = 0.0;
          for (int q = 0; q < D; ++q) {
            const int dy = q / G - md, dx = q % G - md;
            const int yb = y + dy, xb = x + dx;
            if (yb >= 0 && yb < H && xb >= 0 && xb < W)
              a1 += (double)go[((size_t)q * H + y) * W + x] * (double)f2[(size_t)yb * W + xb];
            const int ya = y - dy, xa = x - dx;
            if (ya >= 0 && ya < H && xa >= 0 && xa < W)
              a2 += (double)go[((size_t)q * H + ya) * W + xa] * (double)f1[(size_t)ya * W + xa];
          }
          g1[(size_t)y * W + x] = (float)(a1 / C);
          g2[(size_t)y * W + x] = (float)(a2 / C);
        }
    }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * DeformableConvolution v1 forward (MXNet src/operator/contrib/nn/deformable_im2col.cuh restated).
 * Reference call site: network/layer.py:117-124 with kwargs layer.py:91-95 (3x3, stride 1, dilate 1,
 * pad 1, num_group 1, num_deformable_group 1); offsets built at MaskFlownet.py:230 as 9 copies of
 * (dy, dx) = flow*scale/stride.
 *
 * Sampling rule (border_mode 0, "MXNet-1.5"): a tap at real position (h,w) = (y*s - p + i*d + off_h,
 * x*s - p + j*d + off_w) contributes 0 unless 0 <= h < H and 0 <= w < W.  Inside, the four bilinear
 * corners are floor/floor+1, except that a coordinate whose floor is >= size-1 collapses onto the
 * last row/column with fractional part 0 (no blending with zero).
 * border_mode 1 ("zero-corner", DCNv2 / torchvision rule): contributes unless h <= -1, h >= H,
 * w <= -1 or w >= W; each corner outside the image counts as zero.
 * Offsets channel layout: for deformable group g and tap k = i*kw + j: channel g*2*kh*kw + 2k is the
 * h-offset, +1 the w-offset.   out = W(F x C/groups*kh*kw) . col + bias.
 * ------------------------------------------------------------------------------------------------ */
static inline float sample_tap(const float* plane, int H, int W, float h, float w, int border_mode) {
  if (border_mode == 0) {
    if (!(h >= 0.f && w >= 0.f && h < (float)H && w < (float)W)) return 0.f;
    int h0 = (int)floorf(h), w0 = (int)floorf(w);
    int h1, w1;
    if (h0 >= H - 1) {
      h0 = h1 = H - 1;
      h = (float)h0;
    } else {
      h1 = h0 + 1;
    }
    if (w0 >= W - 1) {
      w0 = w1 = W - 1;
      w = (float)w0;
    } else {
      w1 = w0 + 1;
    }
    const float lh = h - (float)h0, lw = w - (float)w0;
    const float hh = 1.f - lh, hw = 1.f - lw;
    const float v00 = plane[(size_t)h0 * W + w0], v01 = plane[(size_t)h0 * W + w1];
    const float v10 = plane[(size_t)h1 * W + w0], v11 = plane[(size_t)h1 * W + w1];
    return hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11;
  } else {
    if (!(h > -1.f && w > -1.f && h < (float)H && w < (float)W)) return 0.f;
    const int h0 = (int)floorf(h), w0 = (int)floorf(w);
    const int h1 = h0 + 1, w1 = w0 + 1;
    const float lh = h - (float)h0, lw = w - (float)w0;
    const float hh = 1.f - lh, hw = 1.f - lw;
    float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
    if (h0 >= 0 && w0 >= 0) v00 = plane[(size_t)h0 * W + w0];
    if (h0 >= 0 && w1 <= W - 1) v01 = plane[(size_t)h0 * W + w1];
    if (h1 <= H - 1 && w0 >= 0) v10 = plane[(size_t)h1 * W + w0];
    if (h1 <= H - 1 && w1 <= W - 1) v11 = plane[(size_t)h1 * W + w1];
    return hh * hw * v00 + hh * lw * v01 + lh * hw * v10 + lh * lw * v11;
  }
}

MFN_REF_API int mfn_ref_deformable_conv_forward(
    const float* data, const float* offset, const float* weight, const float* bias /* may be NULL */,
    float* out, int N, int C, int H, int W, int F, int kh, int kw, int stride_h, int stride_w,
    int pad_h, int pad_w, int dil_h, int dil_w, int num_group, int num_deformable_group,
    int border_mode, int threads) {
  if (C % num_group || F % num_group || C % num_deformable_group) return -1;
  const int OH = (H + 2 * pad_h - (dil_h * (kh - 1) + 1)) / stride_h + 1;
  const int OW = (W + 2 * pad_w - (dil_w * (kw - 1) + 1)) / stride_w + 1;
  const int Cg = C / num_group, Fg = F / num_group, Cdg = C / num_deformable_group;
  const int K = kh * kw;
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
#endif
  {
    float* col = (float*)malloc(sizeof(float) * (size_t)C * K);
#ifdef _OPENMP
#pragma omp for collapse(2) schedule(static)
#endif
    for (int n = 0; n < N; ++n)
      for (int y = 0; y < OH; ++y)
        for (int x = 0; x < OW; ++x) {
          /* im2col column of this output pixel: col[c*K + k] */
          for (int c = 0; c < C; ++c) {
            const int dg = c / Cdg;
            const float* plane = data + ((size_t)n * C + c) * H * W;
            for (int i = 0; i < kh; ++i)
              for (int j = 0; j < kw; ++j) {
                const int k = i * kw + j;
                const size_t obase = ((size_t)n * num_deformable_group + dg) * 2 * K;
                const float off_h = offset[((obase + 2 * k) * OH + y) * OW + x];
                const float off_w = offset[((obase + 2 * k + 1) * OH + y) * OW + x];
                const float h = (float)(y * stride_h - pad_h + i * dil_h) + off_h;
                const float w = (float)(x * stride_w - pad_w + j * dil_w) + off_w;
                col[(size_t)c * K + k] = sample_tap(plane, H, W, h, w, border_mode);
              }
          }
          for (int f = 0; f < F; ++f) {
            const int g = f / Fg;
            const float* wrow = weight + (size_t)f * Cg * K;
            const float* crow = col + (size_t)g * Cg * K;
            float acc = 0.f;
            for (int t = 0; t < Cg * K; ++t) acc += wrow[t] * crow[t];
            if (bias) acc += bias[f];
            out[(((size_t)n * F + f) * OH + y) * OW + x] = acc;
          }
        }
    free(col);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Upsample(f)  (pure reference code, network/MaskFlownet.py:35-62), restated operation by operation:
 *   pad one replicated row/column at bottom/right (F.pad mode='edge', :51)
 *   transposed convolution, kernel outer(k1,k1), k1[t] = 1 - |c - t|/(c+1), c = f-1, t = 0..2c,
 *   stride f, pad f-1 (:52-60); output rows (H+1-1)*f - 2(f-1) + (2f-1) = f*H + 1
 *   drop the last row and column (:61).
 * Applies per (n,c) plane; `planes` = N*C.
 * ------------------------------------------------------------------------------------------------ */
MFN_REF_API int mfn_ref_upsample(const float* in, float* out, int planes, int H, int W, int factor) {
  if (factor == 1) {
    memcpy(out, in, sizeof(float) * (size_t)planes * H * W);
    return 0;
  }
  const int f = factor, c = f - 1, KS = 2 * f - 1;
  const int PH = H + 1, PW = W + 1;
  const int DH = f * H + 1, DW = f * W + 1; /* deconvolution output before the crop */
  const int OH = f * H, OW = f * W;
  float* k1 = (float*)malloc(sizeof(float) * KS);
  float* padded = (float*)malloc(sizeof(float) * (size_t)PH * PW);
  double* full = (double*)malloc(sizeof(double) * (size_t)DH * DW);
  for (int t = 0; t < KS; ++t) k1[t] = 1.f - fabsf((float)(c - t)) / (float)(c + 1);
  for (int p = 0; p < planes; ++p) {
    const float* src = in + (size_t)p * H * W;
    for (int y = 0; y < PH; ++y)
      for (int x = 0; x < PW; ++x)
        padded[(size_t)y * PW + x] = src[(size_t)(y < H ? y : H - 1) * W + (x < W ? x : W - 1)];
    for (size_t t = 0; t < (size_t)DH * DW; ++t) full[t] = 0.0;
    /* scatter form of a transposed convolution: input (iy,ix) adds k[ky][kx] at
       (iy*f - pad + ky, ix*f - pad + kx) */
    for (int iy = 0; iy < PH; ++iy)
      for (int ix = 0; ix < PW; ++ix) {
        const float v = padded[(size_t)iy * PW + ix];
        for (int ky = 0; ky < KS; ++ky) {
          const int oy = iy * f - c + ky;
          if (oy < 0 || oy >= DH) continue;
          for (int kx = 0; kx < KS; ++kx) {
            const int ox = ix * f - c + kx;
            if (ox < 0 || ox >= DW) continue;
            full[(size_t)oy * DW + ox] += (double)v * (double)(k1[ky] * k1[kx]);
          }
        }
      }
    float* dst = out + (size_t)p * OH * OW;
    for (int y = 0; y < OH; ++y)
      for (int x = 0; x < OW; ++x) dst[(size_t)y * OW + x] = (float)full[(size_t)y * DW + x];
  }
  free(k1);
  free(padded);
  free(full);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * GridGenerator(transform_type='warp') (MXNet src/operator/grid_generator-inl.h restated):
 *   input  flow (N,2,H,W) with channel 0 = x-displacement, channel 1 = y-displacement (the reference
 *          flips its (y,x) flow before the call, layer.py:17)
 *   output grid (N,2,H,W): grid[:,0] = (flow[:,0] + x) / ((W-1)/2) - 1 ; grid[:,1] = (flow[:,1] + y) / ((H-1)/2) - 1
 * ------------------------------------------------------------------------------------------------ */
MFN_REF_API int mfn_ref_grid_generator_warp(const float* flow, float* grid, int N, int H, int W) {
  const float sx = (float)(W - 1) / 2.f, sy = (float)(H - 1) / 2.f;
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t i0 = (((size_t)n * 2 + 0) * H + y) * W + x;
        const size_t i1 = (((size_t)n * 2 + 1) * H + y) * W + x;
        grid[i0] = (flow[i0] + (float)x) / sx - 1.f;
        grid[i1] = (flow[i1] + (float)y) / sy - 1.f;
      }
  return 0;
}

/* BilinearSampler (MXNet src/operator/bilinear_sampler.cc restated): grid in [-1,1] maps to pixel
 * coordinates x = (gx+1)(W-1)/2, y = (gy+1)(H-1)/2; the four neighbours floor/floor+1 contribute
 * their bilinear weight only when they lie inside [0,W-1] x [0,H-1] (zero outside).
 * data (N,C,H,W), grid (N,2,OH,OW) -> out (N,C,OH,OW). */
MFN_REF_API int mfn_ref_bilinear_sampler(const float* data, const float* grid, float* out, int N,
                                         int C, int H, int W, int OH, int OW) {
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < OH; ++y)
      for (int x = 0; x < OW; ++x) {
        const float gx = grid[(((size_t)n * 2 + 0) * OH + y) * OW + x];
        const float gy = grid[(((size_t)n * 2 + 1) * OH + y) * OW + x];
        const float xr = (gx + 1.f) * (float)(W - 1) / 2.f;
        const float yr = (gy + 1.f) * (float)(H - 1) / 2.f;
        const int x0 = (int)floorf(xr), y0 = (int)floorf(yr);
        const float wx0 = 1.f - (xr - (float)x0), wy0 = 1.f - (yr - (float)y0);
        const float wx1 = 1.f - wx0, wy1 = 1.f - wy0;
        const int in00 = (x0 >= 0 && x0 <= W - 1 && y0 >= 0 && y0 <= H - 1);
        const int in01 = (x0 + 1 >= 0 && x0 + 1 <= W - 1 && y0 >= 0 && y0 <= H - 1);
        const int in10 = (x0 >= 0 && x0 <= W - 1 && y0 + 1 >= 0 && y0 + 1 <= H - 1);
        const int in11 = (x0 + 1 >= 0 && x0 + 1 <= W - 1 && y0 + 1 >= 0 && y0 + 1 <= H - 1);
        for (int c = 0; c < C; ++c) {
          const float* p = data + ((size_t)n * C + c) * H * W;
          float v = 0.f;
          if (in00) v += p[(size_t)y0 * W + x0] * wy0 * wx0;
          if (in01) v += p[(size_t)y0 * W + x0 + 1] * wy0 * wx1;
          if (in10) v += p[(size_t)(y0 + 1) * W + x0] * wy1 * wx0;
          if (in11) v += p[(size_t)(y0 + 1) * W + x0 + 1] * wy1 * wx1;
          out[(((size_t)n * C + c) * OH + y) * OW + x] = v;
        }
      }
  return 0;
}
