// augment_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the augmentation kernels (maskflownet_b200/csrc/augment.cu) for the
// host through cuda_shim.h and runs them thread by thread; exported with a C ABI for tests/test_host_logic.py.
//   g++ -O1 -ffp-contract=off -shared -fPIC -DMFN_HOST_EMULATION -I tests/host_emu -x c++ augment_emu.cpp
#define MFN_HOST_EMULATION 1
#include "../../maskflownet_b200/csrc/augment.cu"

using namespace mfn::aug;

template <typename F>
static void for_each_thread(dim3 grid, dim3 block, F body) {
  gridDim = grid;
  blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned t = 0; t < block.x; ++t) {
          blockIdx = dim3(bx, by, bz);
          threadIdx = dim3(t, 0, 0);
          body();
        }
}

extern "C" __attribute__((visibility("default"))) void emu_geometry_augment(
    const void* img1, const void* img2, int is_uint8, const float* flow, const void* mask, int mask_broadcast,
    const float* params, float* o1, float* o2, float* of, float* om, int N, int H, int W, int TH, int TW) {
  const float sx = (float)(2.0 / (double)(TW - 1)), sy = (float)(2.0 / (double)(TH - 1));   // as mfn_geometry_augment_forward
  // two passes: the kernel's per-block table of v / 255 is filled by ALL threads of a block before the barrier; the shim's
  // __shared__ is static and its threads run one after the other, so the first pass fills the table, the second computes
  for (int pass = 0; pass < 2; ++pass) for_each_thread(dim3(3), dim3(64), [&] {
    if (is_uint8)
      geometry_augment_kernel<unsigned char>((const unsigned char*)img1, (const unsigned char*)img2, flow,
                                             (const unsigned char*)mask, mask_broadcast, params, o1, o2, of, om, N, H, W, TH, TW,
                                             sx, sy, 255.f);
    else
      geometry_augment_kernel<float>((const float*)img1, (const float*)img2, flow, (const float*)mask, mask_broadcast, params,
                                     o1, o2, of, om, N, H, W, TH, TW, sx, sy, 1.f);
  });
}

// ws: the per-slice partial sums color_sum_kernel would have written (the test fills them from the oracle's pre-mean image:
// that kernel's block reduction uses warp shuffles, which this shim does not model)
extern "C" __attribute__((visibility("default"))) void emu_color_apply(
    const float* img1, const float* img2, const float* params, const float* noise1, const float* noise2, float sigma,
    long long seed, const float* ws, float* out1, float* out2, int N, int H, int W, int has_pow) {
  // threads 0..2 of a block fill the shared means before the barrier: run every block twice (the shim's __shared__ is static)
  for (int pass = 0; pass < 2; ++pass)
    for (unsigned z = 0; z < 2; ++z)
      for (int n = 0; n < N; ++n) {
        gridDim = dim3(2, N, 2);
        blockDim = dim3(32);
        for (int rep = 0; rep < 2; ++rep)
          for (unsigned bx = 0; bx < 2; ++bx)
            for (unsigned t = 0; t < 32; ++t) {
              blockIdx = dim3(bx, n, z);
              threadIdx = dim3(t, 0, 0);
              color_apply_kernel(img1, img2, noise1, noise2, params, sigma, (unsigned long long)seed, ws, out1, out2, N, H * W,
                                 has_pow);
            }
      }
}

// the pre-mean image (hue / saturation matrix + noise), one value per (n, channel, pixel): what color_sum_kernel sums
extern "C" __attribute__((visibility("default"))) void emu_color_pre_mean(
    const float* img, const float* noise, const float* params, float sigma, long long seed, int image, float* out, int N, int H,
    int W) {
  const int HW = H * W;
  for (int n = 0; n < N; ++n)
    for (int pix = 0; pix < HW; ++pix) {
      float a[3];
      pre_mean(img, noise, params + (size_t)n * COL_P, sigma, (unsigned long long)seed, (unsigned)image, n, HW, pix, a);
      for (int c = 0; c < 3; ++c) out[((size_t)n * 3 + c) * HW + pix] = a[c];
    }
}
