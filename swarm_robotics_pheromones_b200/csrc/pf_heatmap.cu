// pf_heatmap.cu — the reference's offline heat-map loop (Phase-3/graph.py:17-24) on the GPU:
//     heatmap[abs(int(y*H)), abs(int(x*W))] += 20 ; heatmap = gaussian_filter(heatmap, sigma)
// fp64, separable, scipy's default mode='reflect' (d c b a | a b c d), truncate = 4.0.
//
// Bit-compatible with scipy.ndimage.gaussian_filter: the weights are computed on the host exactly as
// scipy's _gaussian_kernel1d does and passed in; each output follows NI_Correlate1D's symmetric
// branch — tmp = x[i]*w[0]; for k = radius..1: tmp += (x[i-k] + x[i+k]) * w[k] — with separately
// rounded adds and multiplies, axis 0 (rows) first, then axis 1.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/pherofield.h"

#define PF_HM_MAX_RADIUS 64

struct pf_heatmap {
  int device, H, W, radius;
  double* buf[2];  // ping-pong
  double* w;       // device weights w[0..radius], w[0] = centre
  char err[256];
};

static char g_hm_err[256] = "";

// scipy 'reflect': index -1 -> 0, -2 -> 1, n -> n-1, n+1 -> n-2 (period 2n)
__device__ __forceinline__ int hm_reflect(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * n;
  i %= p;
  if (i < 0) i += p;
  return i < n ? i : p - 1 - i;
}

// one separable pass; axis 0: neighbours are rows (stride W); axis 1: neighbours are columns
__global__ void __launch_bounds__(256)
k_hm_pass(const double* __restrict__ src, double* __restrict__ dst, int H, int W, int axis, int radius,
          const double* __restrict__ w) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;
  if (col >= W || row >= H) return;
  const int n = axis == 0 ? H : W;
  const int i = axis == 0 ? row : col;
  const long long base = axis == 0 ? (long long)col : (long long)row * W;
  const long long stride = axis == 0 ? W : 1;
  double tmp = __dmul_rn(src[base + (long long)i * stride], w[0]);
  for (int k = radius; k >= 1; --k) {
    const double a = src[base + (long long)hm_reflect(i - k, n) * stride];
    const double b = src[base + (long long)hm_reflect(i + k, n) * stride];
    tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), w[k]));
  }
  dst[(long long)row * W + col] = tmp;
}

__global__ void k_hm_deposit(double* map, int H, int W, double x, double y, double amount) {
  // scale_x = (x / 1) * width ; heatmap[abs(int(scale_y)), abs(int(scale_x))] += 20   (graph.py:19-23)
  long long r = (long long)__dmul_rn(y / 1.0, (double)H);
  long long c = (long long)__dmul_rn(x / 1.0, (double)W);
  r = r < 0 ? -r : r;
  c = c < 0 ? -c : c;
  if (r < H && c < W) map[r * W + c] = __dadd_rn(map[r * W + c], amount);
}

#define HM_CUDA(h, call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      snprintf((h) ? (h)->err : g_hm_err, 256, "%s: %s", #call, cudaGetErrorString(e__));  \
      return e__ == cudaErrorMemoryAllocation ? PF_ENOMEM : PF_ECUDA;                      \
    }                                                                                      \
  } while (0)

extern "C" const char* pf_heatmap_last_error(const pf_heatmap* h) { return h ? h->err : g_hm_err; }

extern "C" int pf_heatmap_destroy(pf_heatmap* h) {
  if (!h) return PF_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (h->buf[0]) cudaFree(h->buf[0]);
  if (h->buf[1]) cudaFree(h->buf[1]);
  if (h->w) cudaFree(h->w);
  delete h;
  return PF_OK;
}

// weights_host: radius+1 doubles, weights_host[0] = centre tap (scipy's normalised kernel is symmetric)
extern "C" int pf_heatmap_create(int device, int H, int W, int radius, const double* weights_host,
                                 pf_heatmap** out) {
  if (!out || H <= 0 || W <= 0 || radius < 0 || radius > PF_HM_MAX_RADIUS || !weights_host) {
    snprintf(g_hm_err, 256, "pf_heatmap_create: bad argument (radius <= %d)", PF_HM_MAX_RADIUS);
    return PF_EINVAL;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    snprintf(g_hm_err, 256, "no CUDA device: pherofield has no CPU fallback");
    return PF_ECUDA;
  }
  pf_heatmap* h = new (std::nothrow) pf_heatmap();
  if (!h) return PF_ENOMEM;
  memset(h, 0, sizeof(*h));
  h->device = device; h->H = H; h->W = W; h->radius = radius;
  const size_t bytes = sizeof(double) * (size_t)H * W;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&h->buf[0], bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->buf[1], bytes);
  if (e == cudaSuccess) e = cudaMalloc(&h->w, sizeof(double) * (radius + 1));
  if (e == cudaSuccess) e = cudaMemset(h->buf[0], 0, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(h->w, weights_host, sizeof(double) * (radius + 1), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    snprintf(g_hm_err, 256, "pf_heatmap_create: %s", cudaGetErrorString(e));
    pf_heatmap_destroy(h);
    return e == cudaErrorMemoryAllocation ? PF_ENOMEM : PF_ECUDA;
  }
  *out = h;
  return PF_OK;
}

// graph.py:19-24 for a batch of trajectory points (host arrays), one deposit + full blur per point
extern "C" int pf_heatmap_accumulate(pf_heatmap* h, int64_t n, const double* x_host, const double* y_host,
                                     double amount, void* stream) {
  if (!h || n < 0 || (n > 0 && (!x_host || !y_host))) return PF_EINVAL;
  HM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid((h->W + 255) / 256, h->H);
  for (int64_t i = 0; i < n; ++i) {
    k_hm_deposit<<<1, 1, 0, s>>>(h->buf[0], h->H, h->W, x_host[i], y_host[i], amount);
    k_hm_pass<<<grid, 256, 0, s>>>(h->buf[0], h->buf[1], h->H, h->W, 0, h->radius, h->w);
    k_hm_pass<<<grid, 256, 0, s>>>(h->buf[1], h->buf[0], h->H, h->W, 1, h->radius, h->w);
  }
  HM_CUDA(h, cudaGetLastError());
  return PF_OK;
}

extern "C" int pf_heatmap_snapshot_host(pf_heatmap* h, double* dst_host, void* stream) {
  if (!h || !dst_host) return PF_EINVAL;
  HM_CUDA(h, cudaSetDevice(h->device));
  HM_CUDA(h, cudaMemcpyAsync(dst_host, h->buf[0], sizeof(double) * (size_t)h->H * h->W, cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
  HM_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
  return PF_OK;
}
