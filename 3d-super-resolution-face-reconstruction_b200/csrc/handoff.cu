// SR -> MICA hand-off on the device. The reference takes every SR image to the host
// (core/metrics.py:16-42 tensor2img), resizes it with OpenCV (model/sr3d/model.py:374 cv2.resize 224x224), builds the
// ArcFace blob with cv2.dnn.blobFromImages (model/sr3d/model.py:127-131) and copies both back to the GPU, one image at
// a time. These kernels produce the same bytes without leaving HBM. The work is HBM-bound integer/byte arithmetic:
// one thread per output element group, coalesced along the innermost dimension, no tensor cores.
//
// Bit-exactness contract (tests/test_gpu_mica_handoff.py, oracle/mica_handoff_oracle.py):
//   * tensor2img: u8 = rint_half_even(((clamp(x,-1,1) + 1) / 2) * 255) with every step rounded to float32
//   * cv::resize INTER_LINEAR, CV_8U: 11-bit fixed-point coefficients, x taps clamped at the border, y rows clipped
//     with the weights kept, vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2
//   * blobFromImages: the 224 -> 112 resize is an exact 2x2 box mean ((a+b+c+d+2)>>2), then (v - 127.5f) * (1/127.5)f in
//     float32, channels 0 and 2 swapped, NCHW
#include <cstdint>

#include "common.cuh"
#include "../../include/b200sr3.h"

namespace b200sr3 {

constexpr int UP = 224, BLOB = 112;

// ---- core/metrics.py:16-42. x: fp32 NCHW [B,C,H,W] -> img: u8 NHWC [B,H,W,C]. A thread owns one pixel (all channels):
// reads are coalesced per channel plane, the C-byte writes of a warp form one contiguous 32*C-byte run.
template <int C>
__global__ void __launch_bounds__(256) tensor2img_kernel(const float* __restrict__ x, uint8_t* __restrict__ img,
                                                         long long npix_total, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // pixel index over [B][H*W]
  if (i >= npix_total) return;
  const long long b = i / HW;
  const int p = (int)(i - b * HW);
  const float* src = x + (size_t)b * C * HW + p;
  uint8_t v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float t = fminf(fmaxf(__ldg(src + (size_t)c * HW), -1.0f), 1.0f);
    t = __fmul_rn(__fadd_rn(t, 1.0f), 0.5f);                 // (t - min) / (max - min): /2 is exact as *0.5
    v[c] = (uint8_t)__float2int_rn(__fmul_rn(t, 255.0f));    // numpy round() = half to even
  }
#pragma unroll
  for (int c = 0; c < C; ++c) img[(size_t)i * C + c] = v[c];
}

// cv::resize coefficient of destination index d along an axis of `sn` source samples (resize.cpp): source index,
// its right/bottom neighbour and the two 11-bit weights.
struct Tap { int s0, s1, a0, a1; };
__device__ __forceinline__ Tap linear_tap(int d, int sn, double scale, bool clamp_weights) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (clamp_weights) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= sn - 1) { f = 0.f; s = sn - 1; }
  }
  Tap t;
  t.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
  t.a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
  t.s0 = min(max(s, 0), sn - 1);
  t.s1 = min(max(s + 1, 0), sn - 1);
  return t;
}

// ---- model/sr3d/model.py:374 + :127-131 + :380-382. A thread owns one pixel of the 112x112 blob = a 2x2 block of
// the 224x224 image: it interpolates the four pixels (3 channels), writes them (u8 NHWC and/or fp32 NCHW / 255) and
// the box-mean blob value of each channel. The source image (<= 48 KB per face at R=128) is read through L1/L2.
__global__ void __launch_bounds__(256) mica_handoff_kernel(const uint8_t* __restrict__ img, int B, int R, double scale,
                                                           uint8_t* __restrict__ up, float* __restrict__ image,
                                                           float* __restrict__ blob) {
  const int bx = blockIdx.x * blockDim.x + threadIdx.x;      // 0 .. 112*112-1
  const int b = blockIdx.y;
  if (bx >= BLOB * BLOB) return;
  const int oy = bx / BLOB, ox = bx - oy * BLOB;
  const uint8_t* src = img + (size_t)b * R * R * 3;
  Tap tx[2], ty[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    tx[k] = linear_tap(2 * ox + k, R, scale, true);
    ty[k] = linear_tap(2 * oy + k, R, scale, false);
  }
  int sum[3] = {0, 0, 0};
#pragma unroll
  for (int ky = 0; ky < 2; ++ky) {
    const uint8_t* r0 = src + ((size_t)ty[ky].s0 * R) * 3;
    const uint8_t* r1 = src + ((size_t)ty[ky].s1 * R) * 3;
    const int dy = 2 * oy + ky;
    int v[2][3];
#pragma unroll
    for (int kx = 0; kx < 2; ++kx) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = (int)r0[tx[kx].s0 * 3 + c] * tx[kx].a0 + (int)r0[tx[kx].s1 * 3 + c] * tx[kx].a1;
        const int h1 = (int)r1[tx[kx].s0 * 3 + c] * tx[kx].a0 + (int)r1[tx[kx].s1 * 3 + c] * tx[kx].a1;
        v[kx][c] = (((ty[ky].a0 * (h0 >> 4)) >> 16) + ((ty[ky].a1 * (h1 >> 4)) >> 16) + 2) >> 2;
        sum[c] += v[kx][c];
      }
    }
    // the thread's two pixels of this row are adjacent: 6 bytes (2-byte aligned) / one float2 per channel plane
    if (up) {
      unsigned short* d = reinterpret_cast<unsigned short*>(up + (((size_t)b * UP + dy) * UP + 2 * ox) * 3);
      d[0] = (unsigned short)(v[0][0] | (v[0][1] << 8));
      d[1] = (unsigned short)(v[0][2] | (v[1][0] << 8));
      d[2] = (unsigned short)(v[1][1] | (v[1][2] << 8));
    }
    if (image) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        *reinterpret_cast<float2*>(image + (((size_t)b * 3 + c) * UP + dy) * UP + 2 * ox) =
            make_float2(__fdiv_rn((float)v[0][c], 255.0f), __fdiv_rn((float)v[1][c], 255.0f));   // == float(v / 255.) for all 256 values
    }
  }
  if (blob) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float small = (float)((sum[c] + 2) >> 2);
      const float v = __fmul_rn(__fsub_rn(small, 127.5f), (float)(1.0 / 127.5));
      blob[(((size_t)b * 3 + (2 - c)) * BLOB + oy) * BLOB + ox] = v;      // swapRB
    }
  }
}

// ---- model3 path: core/metrics.py:44-50 tensor2tensor_img(x) * 255 -> model/sr3d/model.py:105-124 create_tensor_blob
// ((v - 127.5) / 127.5 -> F.interpolate(bilinear, align_corners=False, 112x112) -> channels swapped). Float path:
// parity within 1e-5 (the interpolation weights differ from torch's by an ulp of the source index).
__global__ void __launch_bounds__(256) tensor_blob_kernel(const float* __restrict__ x, int R, float scale,
                                                          float* __restrict__ blob) {
  const int bx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (bx >= BLOB * BLOB) return;
  const int oy = bx / BLOB, ox = bx - oy * BLOB;
  auto tap = [&](int d, int& i0, int& i1, float& l0, float& l1) {
    const float s = fmaxf(__fsub_rn(__fmul_rn((float)d + 0.5f, scale), 0.5f), 0.f);
    i0 = min((int)s, R - 1);
    i1 = min(i0 + 1, R - 1);
    l1 = s - (float)i0;
    l0 = 1.0f - l1;
  };
  int x0, x1, y0, y1;
  float hx0, hx1, hy0, hy1;
  tap(ox, x0, x1, hx0, hx1);
  tap(oy, y0, y1, hy0, hy1);
  auto val = [&](const float* plane, int yy, int xx) {
    float t = fminf(fmaxf(__ldg(plane + (size_t)yy * R + xx), -1.0f), 1.0f);
    t = __fmul_rn(__fadd_rn(t, 1.0f), 0.5f);
    return __fdiv_rn(__fsub_rn(__fmul_rn(t, 255.0f), 127.5f), 127.5f);
  };
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* plane = x + ((size_t)b * 3 + c) * R * R;
    const float top = __fadd_rn(__fmul_rn(hx0, val(plane, y0, x0)), __fmul_rn(hx1, val(plane, y0, x1)));
    const float bot = __fadd_rn(__fmul_rn(hx0, val(plane, y1, x0)), __fmul_rn(hx1, val(plane, y1, x1)));
    blob[(((size_t)b * 3 + (2 - c)) * BLOB + oy) * BLOB + ox] = __fadd_rn(__fmul_rn(hy0, top), __fmul_rn(hy1, bot));
  }
}

static void require_device() {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(std::string("no CUDA device available (") + cudaGetErrorString(e) + "): no CPU fallback");
}

void launch_tensor2img(const float* x, int B, int C, int H, int W, uint8_t* img, cudaStream_t s) {
  REQUIRE(x && img, "tensor2img: null pointer");
  REQUIRE(B >= 1 && H >= 1 && W >= 1 && (C == 1 || C == 3), "tensor2img: expects [B,1|3,H,W]");
  require_device();
  const long long n = (long long)B * H * W;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (C == 3) tensor2img_kernel<3><<<blocks, 256, 0, s>>>(x, img, n, H * W);
  else tensor2img_kernel<1><<<blocks, 256, 0, s>>>(x, img, n, H * W);
  CUDA_CHECK(cudaGetLastError());
}

void launch_mica_handoff(const uint8_t* img, int B, int R, uint8_t* up, float* image, float* blob, cudaStream_t s) {
  REQUIRE(img && (up || image || blob), "mica_handoff: null pointer");
  REQUIRE(B >= 1 && B <= 65535 && R >= 2 && R <= 4096, "mica_handoff: expects 1 <= B <= 65535 square images of side 2..4096");
  REQUIRE(R != 2 * UP, "mica_handoff: a 2x down-scale takes OpenCV's INTER_AREA path, which is not implemented");
  require_device();
  const double scale = 1.0 / ((double)UP / (double)R);       // cv::resize: scale = 1 / (dsize / ssize)
  dim3 grid((BLOB * BLOB + 255) / 256, B);
  mica_handoff_kernel<<<grid, 256, 0, s>>>(img, B, R, scale, up, image, blob);
  CUDA_CHECK(cudaGetLastError());
}

void launch_tensor_blob(const float* x, int B, int R, float* blob, cudaStream_t s) {
  REQUIRE(x && blob, "tensor_blob: null pointer");
  REQUIRE(B >= 1 && B <= 65535 && R >= 1, "tensor_blob: expects [B,3,R,R], B <= 65535");
  require_device();
  dim3 grid((BLOB * BLOB + 255) / 256, B);
  tensor_blob_kernel<<<grid, 256, 0, s>>>(x, R, (float)R / (float)BLOB, blob);
  CUDA_CHECK(cudaGetLastError());
}

}  // namespace b200sr3
