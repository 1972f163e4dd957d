// synth.cu -- deterministic synthetic terrain for benchmarks and tests (not reference API).
// fBm-like field: 12 octaves of smoothstep-interpolated lattice value noise, wavelength
// 4096 .. 2 cells, amplitude ~ wavelength^0.8 (Hurst 0.8).  A pure function of
// (seed, global row, global col), so row stripes generated on different GPUs tile exactly.
#include "common.cuh"

namespace xrs {

__device__ __forceinline__ uint32_t hash3(uint32_t x, uint32_t y, uint32_t s) {
    uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u) ^ (s * 0xC2B2AE3Du);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
__device__ __forceinline__ float lattice(uint32_t ix, uint32_t iy, uint32_t s) {
    return (float)(hash3(ix, iy, s) >> 8) * (1.0f / 8388608.0f) - 1.0f;  // [-1, 1)
}

__global__ void __launch_bounds__(256) synth_kernel(float *out, int64_t pitch_elems, int64_t H, int64_t W,
                                                    int64_t row0, int64_t col0, uint32_t seed, float zmin,
                                                    float zmax) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t y = blockIdx.y;
    if (x >= W || y >= H) return;
    const int64_t gx = x + col0, gy = y + row0;
    float acc = 0.f, norm = 0.f;
#pragma unroll 1
    for (int o = 0; o < 12; ++o) {
        const int shift = 12 - o;  // wavelength 2^shift cells
        const float amp = exp2f(0.8f * (float)(shift - 12));
        const int64_t cx = gx >> shift, cy = gy >> shift;
        const float inv = 1.0f / (float)(1 << shift);
        float fx = (float)(gx - (cx << shift)) * inv, fy = (float)(gy - (cy << shift)) * inv;
        fx = fx * fx * (3.f - 2.f * fx);
        fy = fy * fy * (3.f - 2.f * fy);
        const uint32_t s = seed + 0x632BE5ABu * (uint32_t)o;
        const float v00 = lattice((uint32_t)cx, (uint32_t)cy, s), v10 = lattice((uint32_t)cx + 1, (uint32_t)cy, s);
        const float v01 = lattice((uint32_t)cx, (uint32_t)cy + 1, s), v11 = lattice((uint32_t)cx + 1, (uint32_t)cy + 1, s);
        const float a = v00 + (v10 - v00) * fx, b = v01 + (v11 - v01) * fx;
        acc += amp * (a + (b - a) * fy);
        norm += amp;
    }
    float t = 0.5f + 0.5f * acc / norm * 1.8f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    out[y * pitch_elems + x] = zmin + (zmax - zmin) * t;
}

}  // namespace xrs

extern "C" int xrs_synth_terrain_f32(float *out, int64_t out_pitch, int64_t H, int64_t W, int64_t row0,
                                     int64_t col0, uint64_t seed, float zmin, float zmax, xrs_stream_t s) {
    if (H <= 0 || W <= 0) return XRS_OK;
    XRS_REQUIRE(out != nullptr && out_pitch % 4 == 0 && out_pitch >= W * 4, "bad output / pitch");
    XRS_REQUIRE(H <= 65535LL * 1024, "too many rows for one launch");
    // rows go to gridDim.y in chunks of 65535
    for (int64_t y0 = 0; y0 < H; y0 += 65535) {
        const int64_t h = (H - y0) < 65535 ? (H - y0) : 65535;
        dim3 grid((unsigned)((W + 255) / 256), (unsigned)h);
        xrs::synth_kernel<<<grid, 256, 0, (cudaStream_t)s>>>(out + y0 * (out_pitch / 4), out_pitch / 4, h, W,
                                                            row0 + y0, col0, (uint32_t)(seed * 0x9E3779B97F4A7C15ull >> 32),
                                                            zmin, zmax);
    }
    XRS_CUDA(cudaGetLastError());
    return XRS_OK;
}
