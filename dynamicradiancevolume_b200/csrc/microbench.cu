// microbench.cu — measured roofline denominators for the gather / apply
// kernels ("FP32 peak must be measured on the box with an FFMA
// micro-kernel"; MEASURED_PEAKS.json only carries HBM copy and bf16 GEMM).
// Each benchmark runs a resident grid for a fixed iteration count and reports
// the best of 5 timed launches (CUDA events).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/drv_gi.h"

namespace {

constexpr int kIters = 4096;

// 16 independent FFMA chains per thread, 3 distinct register sources each.
__global__ void __launch_bounds__(256) ffma_kernel(float* out, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  float x = a + threadIdx.x * 1e-9f, y = b;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Same with packed FP32x2 FMAs (FFMA2): 16 float2 chains.
__global__ void __launch_bounds__(256) ffma2_kernel(float* out, float a, float b) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2((float)(threadIdx.x + i), (float)i);
  float2 x = make_float2(a + threadIdx.x * 1e-9f, a), y = make_float2(b, b + 1e-9f);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The accumulate pattern of the gather: acc_k += f * u_k with 12 accumulators
// and per-iteration varying multipliers (register-bank pressure like the real loop).
__global__ void __launch_bounds__(256) ffma_outer_kernel(float* out, float a) {
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  float f0 = a, f1 = a * 1.1f, f2 = a * 1.2f, u0 = 0.3f + threadIdx.x * 1e-7f, u1 = 0.4f, u2 = 0.5f, u3 = 0.6f;
  for (int it = 0; it < kIters; ++it) {
    acc[0] = fmaf(f0, u0, acc[0]); acc[1] = fmaf(f1, u0, acc[1]); acc[2] = fmaf(f2, u0, acc[2]);
    acc[3] = fmaf(f0, u1, acc[3]); acc[4] = fmaf(f1, u1, acc[4]); acc[5] = fmaf(f2, u1, acc[5]);
    acc[6] = fmaf(f0, u2, acc[6]); acc[7] = fmaf(f1, u2, acc[7]); acc[8] = fmaf(f2, u2, acc[8]);
    acc[9] = fmaf(f0, u3, acc[9]); acc[10] = fmaf(f1, u3, acc[10]); acc[11] = fmaf(f2, u3, acc[11]);
    u0 += 1e-6f; u1 += 1e-6f; u2 += 1e-6f; u3 += 1e-6f;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) mufu_kernel(float* out, float a) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = a + (float)(threadIdx.x + i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float r;
      asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i]));
      acc[i] = r;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Broadcast 128-bit shared loads (every lane reads the same address), as the VPL fetch does.
__global__ void __launch_bounds__(256) lds_broadcast_kernel(float* out) {
  __shared__ float4 s[512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s[i] = make_float4((float)i, 1.f, 2.f, 3.f);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = 0; it < kIters / 8; ++it) {
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      float4 v = s[(it + i * 8) & 511];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// L2-resident streaming read (buffer of 48 MB << 126 MB L2), 128-bit loads.
__global__ void __launch_bounds__(256) l2_read_kernel(const uint4* __restrict__ buf, size_t n16, int reps, uint32_t* out) {
  uint32_t acc = 0;
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = tid; i < n16; i += stride) {
      uint4 v = __ldcg(buf + i);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  if (acc == 0xdeadbeefu) out[0] = acc;
}

// Legacy warp-level tensor-core path: mma.sync.m16n8k8 TF32 with FP32 accumulate, 4 independent accumulator
// sets per warp (the gather's mma variant feeds its accumulate step through this instruction).
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void __launch_bounds__(256) mma_tf32_kernel(float* out, float a) {
  float d[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t A[4] = {__float_as_uint(a), __float_as_uint(a * 1.5f), __float_as_uint(a + threadIdx.x * 1e-3f), __float_as_uint(a * 0.5f)};
  uint32_t B[2] = {__float_as_uint(0.25f), __float_as_uint(0.5f + threadIdx.x * 1e-3f)};
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) mma_tf32(d[i], A, B);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// The mma gather's instruction mix per pair-warp: 10 FFMA2 + 2 mma.sync (operands produced by the FFMA2 chain).
__global__ void __launch_bounds__(256) mma_mix_kernel(float* out, float a) {
  float d[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  float2 x[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) x[i] = make_float2(a + i, a - i + threadIdx.x * 1e-6f);
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-6f, -1e-6f);
  uint32_t B[2] = {__float_as_uint(0.25f), __float_as_uint(0.5f + threadIdx.x * 1e-3f)};
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 10; ++i) x[i] = __ffma2_rn(x[i], m, c);
    uint32_t A0[4] = {__float_as_uint(x[0].x), __float_as_uint(x[1].x), __float_as_uint(x[0].y), __float_as_uint(x[1].y)};
    uint32_t A1[4] = {__float_as_uint(x[2].x), __float_as_uint(x[3].x), __float_as_uint(x[2].y), __float_as_uint(x[3].y)};
    mma_tf32(d[0], A0, B);
    mma_tf32(d[1], A1, B);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
#pragma unroll
  for (int i = 0; i < 10; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Operand-delivery study of the gather's accumulate step: acc[m][c] += f[c] * u[m] over NU basis values and 3
// colours, packed (FFMA2) or scalar, issued colour-major (consecutive FMAs share f[c]) or basis-major (share
// u[m]). Every FMA reads three registers, one of them a private accumulator; what fraction of the FP32 pipe's
// lane-cycles such a stream sustains is what bounds the pair loop.
template <int NU, bool FMAJOR, bool PACKED>
__global__ void __launch_bounds__(256) outer_study_kernel(float* out, float a) {
  float2 acc[NU][3];
#pragma unroll
  for (int m = 0; m < NU; ++m)
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[m][c] = make_float2(0.f, 0.f);
  float2 f[3] = {make_float2(a, a * 1.01f), make_float2(a * 1.1f, a * 1.11f), make_float2(a * 1.2f, a * 1.21f)};
  float2 u[NU];
#pragma unroll
  for (int m = 0; m < NU; ++m) u[m] = make_float2(0.3f + 0.1f * m + threadIdx.x * 1e-7f, 0.35f + 0.1f * m);
  const float2 du = make_float2(1e-6f, 2e-6f);
  for (int it = 0; it < kIters; ++it) {
    if (PACKED) {
      if (FMAJOR) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int m = 0; m < NU; ++m) acc[m][c] = __ffma2_rn(f[c], u[m], acc[m][c]);
      } else {
#pragma unroll
        for (int m = 0; m < NU; ++m)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[m][c] = __ffma2_rn(f[c], u[m], acc[m][c]);
      }
#pragma unroll
      for (int m = 0; m < NU; ++m) u[m] = __fadd2_rn(u[m], du);
    } else {
      if (FMAJOR) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int m = 0; m < NU; ++m) acc[m][c].x = fmaf(f[c].x, u[m].x, acc[m][c].x);
      } else {
#pragma unroll
        for (int m = 0; m < NU; ++m)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[m][c].x = fmaf(f[c].x, u[m].x, acc[m][c].x);
      }
#pragma unroll
      for (int m = 0; m < NU; ++m) u[m].x += du.x;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < NU; ++m)
#pragma unroll
    for (int c = 0; c < 3; ++c) s += acc[m][c].x + acc[m][c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~Timer() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

template <typename F>
double best_ms(F launch) {
  Timer t;
  launch();
  cudaDeviceSynchronize();
  double best = 1e30;
  for (int i = 0; i < 5; ++i) {
    cudaEventRecord(t.a);
    launch();
    cudaEventRecord(t.b);
    cudaEventSynchronize(t.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.a, t.b);
    if (ms < best) best = ms;
  }
  return best;
}

const char* kNames[] = {
    "ffma_tflops",         // scalar FFMA, 3 register sources: TFLOP/s (FMA = 2 flop)
    "ffma2_tflops",        // packed FFMA2: TFLOP/s
    "ffma_outer_tflops",   // the gather's accumulate pattern: TFLOP/s
    "mufu_rsq_gops",       // MUFU.RSQ: G ops/s
    "lds128_bcast_tbs",    // broadcast LDS.128: TB/s of register fill (16 B x 32 lanes per instruction)
    "l2_read_gbs",         // L2-resident 128-bit loads: GB/s
    "sm_clock_mhz",        // current SM clock reported by the driver
    "mma_tf32_tflops",     // mma.sync.m16n8k8 TF32 (legacy tensor path): TFLOP/s (2*16*8*8 flop per instruction)
    "mma_mix_cyc",         // 10 FFMA2 + 2 mma.sync per iteration: SM cycles per iteration per SM sub-partition warp slot
    // operand-delivery study (outer_study_kernel): fraction of the FP32 pipe's lane-cycles at the current clock;
    // name = <packed|scalar>_<colour|basis>major_nu<4|8>_w<warps per SM>
    "study_scalar_cmajor_nu4_w8", "study_scalar_cmajor_nu4_w32", "study_scalar_bmajor_nu4_w32",
    "study_packed_cmajor_nu4_w8", "study_packed_cmajor_nu4_w32", "study_packed_bmajor_nu4_w8", "study_packed_bmajor_nu4_w32",
    "study_packed_cmajor_nu8_w8", "study_packed_cmajor_nu8_w16", "study_packed_bmajor_nu8_w8", "study_packed_bmajor_nu8_w16",
};

} // namespace

extern "C" uint32_t drv_microbench_count(void) { return (uint32_t)(sizeof(kNames) / sizeof(kNames[0])); }
extern "C" const char* drv_microbench_name(uint32_t which) { return which < drv_microbench_count() ? kNames[which] : "?"; }

extern "C" drv_status drv_microbench(int32_t device, uint32_t which, double* result) {
  if (!result || which >= drv_microbench_count()) return DRV_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return DRV_ERR_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return DRV_ERR_CUDA;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256;
  float* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return DRV_ERR_CUDA;
  const double lanes = (double)blocks * threads;
  double r = 0.0;
  switch (which) {
    case 0: { double ms = best_ms([&] { ffma_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
              r = lanes * kIters * 16 * 2.0 / (ms * 1e-3) / 1e12; } break;
    case 1: { double ms = best_ms([&] { ffma2_kernel<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
              r = lanes * kIters * 16 * 4.0 / (ms * 1e-3) / 1e12; } break;
    case 2: { double ms = best_ms([&] { ffma_outer_kernel<<<blocks, threads>>>(out, 1.0001f); });
              r = lanes * kIters * 12 * 2.0 / (ms * 1e-3) / 1e12; } break; // the 4 FADDs are not counted
    case 3: { double ms = best_ms([&] { mufu_kernel<<<blocks, threads>>>(out, 1.5f); });
              r = lanes * kIters * 8 / (ms * 1e-3) / 1e9; } break;
    case 4: { double ms = best_ms([&] { lds_broadcast_kernel<<<blocks, threads>>>(out); });
              r = lanes * (kIters / 8) * 64 * 16.0 / (ms * 1e-3) / 1e12; } break;
    case 5: {
      const size_t bytes = 48ull << 20;
      uint4* buf = nullptr;
      uint32_t* o2 = nullptr;
      if (cudaMalloc(&buf, bytes) != cudaSuccess) { cudaFree(out); return DRV_ERR_CUDA; }
      cudaMalloc(&o2, 4);
      cudaMemset(buf, 1, bytes);
      const int reps = 8;
      double ms = best_ms([&] { l2_read_kernel<<<sms * 8, 256>>>(buf, bytes / 16, reps, o2); });
      r = (double)bytes * reps / (ms * 1e-3) / 1e9;
      cudaFree(buf); cudaFree(o2);
    } break;
    case 7: { double ms = best_ms([&] { mma_tf32_kernel<<<blocks, threads>>>(out, 1.0001f); });
              r = (lanes / 32.0) * kIters * 4 * 2048.0 / (ms * 1e-3) / 1e12; } break;
    case 8: { double ms = best_ms([&] { mma_mix_kernel<<<blocks, threads>>>(out, 1.0001f); });
              int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
              // warps per SM sub-partition = blocks/sms * threads/32 / 4; cycles per iteration per warp slot
              r = (ms * 1e-3) * (khz * 1e3) / kIters / ((double)blocks / sms * threads / 32.0 / 4.0); } break;
    case 6: { int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device); r = khz / 1000.0; } break;
    default: {
      int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
      struct Study { int nu; bool cmajor, packed; int warps; };
      static const Study st[] = {{4, true, false, 8}, {4, true, false, 32}, {4, false, false, 32},
                                 {4, true, true, 8}, {4, true, true, 32}, {4, false, true, 8}, {4, false, true, 32},
                                 {8, true, true, 8}, {8, true, true, 16}, {8, false, true, 8}, {8, false, true, 16}};
      const Study& S = st[which - 9];
      const int b = sms * S.warps / 8;
      double ms = 0;
#define DRV_ST(NU, CM, PK) ms = best_ms([&] { outer_study_kernel<NU, CM, PK><<<b, threads>>>(out, 1.0001f); })
      if (S.nu == 4) { if (S.packed) { if (S.cmajor) DRV_ST(4, true, true); else DRV_ST(4, false, true); }
                       else { if (S.cmajor) DRV_ST(4, true, false); else DRV_ST(4, false, false); } }
      else { if (S.cmajor) DRV_ST(8, true, true); else DRV_ST(8, false, true); }
#undef DRV_ST
      const double lane_cycles = (double)b * threads * kIters * (4.0 * S.nu) * (S.packed ? 2.0 : 1.0);
      r = lane_cycles / (ms * 1e-3) / ((double)sms * 128.0 * khz * 1e3);
    } break;
  }
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return DRV_ERR_CUDA;
  *result = r;
  return DRV_OK;
}
