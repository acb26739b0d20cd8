// loss_emu.cpp -- TEST INFRASTRUCTURE ONLY.  Compiles the fused MultiscaleEpe kernels (maskflownet_b200/csrc/loss.cu) for the
// host through cuda_shim.h (one lane per "warp", one thread per block: the reductions degenerate to their serial parts) and
// runs them thread by thread; C ABI for tests/test_train_augment.py.
//   g++ -O1 -ffp-contract=off -shared -fPIC -I tests/host_emu loss_emu.cpp
#define MFN_HOST_EMULATION 1
#include "../../maskflownet_b200/csrc/loss.cu"

using namespace mfn::epe;

static Args make_args(const float* const* preds, float* const* gpreds, const int* scales, const float* weights, int num, int N,
                      int H, int W) {
  Args A;
  A.num = num;
  A.first_warp[0] = 0;
  for (int s = 0; s < num; ++s) {
    A.pred[s] = preds[s];
    A.gpred[s] = gpreds ? gpreds[s] : nullptr;
    A.scale[s] = scales[s];
    A.weight[s] = weights[s];
    A.first_warp[s + 1] = A.first_warp[s] + (long long)N * (H / scales[s]) * (W / scales[s]);
  }
  return A;
}

extern "C" __attribute__((visibility("default"))) void emu_epe_forward(
    const float* flow, const float* mask, const float* const* preds, const int* scales, const float* weights, int num, float eps,
    float q, float* loss, float* mask_sum, int N, int H, int W) {
  const Args A = make_args(preds, nullptr, scales, weights, num, N, H, W);
  const int blocks = 5;
  float* partial = new float[(size_t)N * blocks * 2];
  gridDim = dim3(blocks, N);
  blockDim = dim3(1);
  threadIdx = dim3(0);
  for (int n = 0; n < N; ++n)
    for (int b = 0; b < blocks; ++b) {
      blockIdx = dim3(b, n);
      epe_forward_kernel(flow, mask, A, eps, q, partial, H, W);
    }
  gridDim = dim3(1);
  blockDim = dim3(N);
  blockIdx = dim3(0);
  for (int n = 0; n < N; ++n) {
    threadIdx = dim3(n);
    epe_finish_kernel(partial, loss, mask_sum, N, blocks);
  }
  delete[] partial;
}

extern "C" __attribute__((visibility("default"))) void emu_epe_backward(
    const float* flow, const float* mask, const float* const* preds, const int* scales, const float* weights, int num, float eps,
    float q, const float* grad_loss, const float* mask_sum, float* const* gpreds, int N, int H, int W) {
  const Args A = make_args(preds, gpreds, scales, weights, num, N, H, W);
  gridDim = dim3(7);          // fewer "warps" than work items: the grid-stride loop is exercised
  blockDim = dim3(3);
  for (unsigned b = 0; b < 7; ++b)
    for (unsigned t = 0; t < 3; ++t) {
      blockIdx = dim3(b);
      threadIdx = dim3(t);
      epe_backward_kernel(flow, mask, A, eps, q, grad_loss, mask_sum, N, H, W);
    }
}
