// Layer 0 of the network in ONE kernel: Focus space-to-depth (reference models/common.py:549-550) + 3x3 Conv(12 -> Co) with folded BN +
// SiLU, straight from the caller's NCHW image (fp32 / fp16 / uint8 scaled by 1/255 like detect.py:137) to the NHWC fp16 activation.
//
// Why not the tcgen05 kernel: with only 12 input channels an implicit-GEMM row is 32 bytes, and TMA issues one request per row -
// the layer ends up bound by the TMA row rate, 4x above its HBM roofline (profiles/README.md).  Here a CTA converts the raw halo
// (20 x 132 x 3 pixels for an 8 x 64 output tile) ONCE into a space-to-depth fp16 tile in shared memory and feeds legacy
// mma.sync.m16n8k16 (K = 9 taps x 16 padded channels) from it with ldmatrix; the layer is memory bound (14.5 GFLOP vs 201 MB per
// batch-16 step), so HMMA throughput is irrelevant and the input is read from HBM exactly once.
#include "kernels.h"

namespace myolo {

static constexpr int kFcTileH = 8, kFcTileW = 64;            // output pixels per CTA: 8 rows (one per warp) x 64 columns
static constexpr int kFcHaloW = kFcTileW + 2, kFcHaloH = kFcTileH + 2;

template <typename T> __device__ __forceinline__ float fc_unit(T v);
template <> __device__ __forceinline__ float fc_unit<float>(float v) { return v; }
template <> __device__ __forceinline__ float fc_unit<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float fc_unit<uint8_t>(uint8_t v) { return (float)v / 255.0f; }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NT = Co / 8 (4 for yolov5s: Co 32, 6 for yolov5m: Co 48)
template <typename T, int NT>
__global__ void __launch_bounds__(256) focus_conv_kernel(const T* __restrict__ x, int B, int H, int W, const __half* __restrict__ wp,
                                                         const float* __restrict__ bias, TensorView out) {
  // s2d halo tile [10][66][16 halves] (32-byte rows), weights [Co][144 halves]
  __shared__ __align__(16) __half s_in[kFcHaloH * kFcHaloW * 16];
  __shared__ __align__(16) __half s_w[NT * 8 * 144];
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = (Wo + kFcTileW - 1) / kFcTileW;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x, b = blockIdx.y;
  const int ox0 = tx * kFcTileW, oy0 = ty * kFcTileH;

  for (int i = threadIdx.x; i < NT * 8 * 144 / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(wp) + i);
  // space-to-depth conversion of the halo: channel = g*3 + c, g = (dy,dx) in order (0,0),(1,0),(0,1),(1,1); 12..15 zero
  for (int i = threadIdx.x; i < kFcHaloH * kFcHaloW; i += blockDim.x) {
    const int hy = i / kFcHaloW, hx = i - hy * kFcHaloW;
    const int sy = oy0 - 1 + hy, sx = ox0 - 1 + hx;
    __align__(16) __half v[16];
    if (sy >= 0 && sy < Ho && sx >= 0 && sx < Wo) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const T* p = x + (((size_t)b * 3 + c) * H + 2 * sy) * W + 2 * sx;
        v[0 * 3 + c] = __float2half_rn(fc_unit<T>(p[0]));
        v[2 * 3 + c] = __float2half_rn(fc_unit<T>(p[1]));
        v[1 * 3 + c] = __float2half_rn(fc_unit<T>(p[W]));
        v[3 * 3 + c] = __float2half_rn(fc_unit<T>(p[W + 1]));
      }
#pragma unroll
      for (int c = 12; c < 16; ++c) v[c] = __float2half_rn(0.f);
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = __float2half_rn(0.f);
    }
    uint4* d = reinterpret_cast<uint4*>(s_in + (size_t)i * 16);
    d[0] = reinterpret_cast<uint4*>(v)[0];
    d[1] = reinterpret_cast<uint4*>(v)[1];
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // warp = output row inside the tile
  float acc[4][NT][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][n][e] = 0.f;

  const uint32_t s_in_u = smem_u32(s_in), s_w_u = smem_u32(s_w);
  // ldmatrix row providers: A: lane -> pixel (lane & 15), k half (lane >> 4); B: lane -> n row (lane & 7) + 8*((lane >> 4) & 1), k half ((lane >> 3) & 1)
  const int a_pix = lane & 15, a_kh = lane >> 4;
  const int b_n = (lane & 7) + ((lane >> 4) << 3), b_kh = (lane >> 3) & 1;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int ky = t / 3, kx = t % 3;
    uint32_t bf[NT][2];
#pragma unroll
    for (int n2 = 0; n2 < NT / 2; ++n2) {
      const uint32_t addr = s_w_u + (((n2 * 16 + b_n) * 144) + t * 16 + b_kh * 8) * 2;
      ldmatrix_x4(addr, bf[2 * n2][0], bf[2 * n2][1], bf[2 * n2 + 1][0], bf[2 * n2 + 1][1]);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int hx = m * 16 + a_pix + kx, hy = warp + ky;
      const uint32_t addr = s_in_u + ((hy * kFcHaloW + hx) * 16 + a_kh * 8) * 2;
      uint32_t a0, a1, a2, a3;
      ldmatrix_x4(addr, a0, a1, a2, a3);
#pragma unroll
      for (int n = 0; n < NT; ++n) mma_16816(acc[m][n], a0, a1, a2, a3, bf[n][0], bf[n][1]);
    }
  }

  // epilogue: bias + SiLU -> fp16 NHWC.  C fragment: rows lane/4 and lane/4+8, channels 8n + 2*(lane%4) + {0,1}
  const int oy = oy0 + warp;
  if (oy >= Ho) return;
  const int cq = (lane & 3) * 2, rq = lane >> 2;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int ox = ox0 + m * 16 + rq + hrow * 8;
      if (ox >= Wo) continue;
      __half* op = reinterpret_cast<__half*>(out.base) + (((size_t)b * Ho + oy) * Wo + ox) * out.ctot;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int ch = n * 8 + cq;
        const float v0 = silu_f(acc[m][n][hrow * 2 + 0] + __ldg(bias + ch));
        const float v1 = silu_f(acc[m][n][hrow * 2 + 1] + __ldg(bias + ch + 1));
        *reinterpret_cast<__half2*>(op + ch) = __floats2half2_rn(v0, v1);
      }
    }
  }
}

int launch_focus_conv(const void* x, int x_dtype, int B, int H, int W, const __half* wp, const float* bias, int co, const TensorView& out,
                      cudaStream_t s) {
  MYOLO_REQUIRE(H % 2 == 0 && W % 2 == 0 && out.H == H / 2 && out.W == W / 2 && out.dtype == MYOLO_F16 && out.C == co,
                "focus_conv: bad output view");
  MYOLO_REQUIRE(co == 32 || co == 48, "focus_conv: Co must be 32 or 48 (got %d)", co);
  const dim3 grid(ceil_div(W / 2, kFcTileW) * ceil_div(H / 2, kFcTileH), B);
#define FC_LAUNCH(T, NT) focus_conv_kernel<T, NT><<<grid, 256, 0, s>>>((const T*)x, B, H, W, wp, bias, out)
  if (co == 32) {
    if (x_dtype == MYOLO_F32) FC_LAUNCH(float, 4);
    else if (x_dtype == MYOLO_F16) FC_LAUNCH(__half, 4);
    else if (x_dtype == MYOLO_U8) FC_LAUNCH(uint8_t, 4);
    else MYOLO_REQUIRE(false, "focus_conv: unsupported dtype %d", x_dtype);
  } else {
    if (x_dtype == MYOLO_F32) FC_LAUNCH(float, 6);
    else if (x_dtype == MYOLO_F16) FC_LAUNCH(__half, 6);
    else if (x_dtype == MYOLO_U8) FC_LAUNCH(uint8_t, 6);
    else MYOLO_REQUIRE(false, "focus_conv: unsupported dtype %d", x_dtype);
  }
#undef FC_LAUNCH
  MYOLO_LAUNCH_CHECK();
  return 0;
}

}  // namespace myolo
