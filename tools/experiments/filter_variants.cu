// Steady-state FMA throughput of the membership filter's inner loop, by instruction form and
// operand order (no refill, no decisions: the (proposal slots) x (tile) product and the
// per-group maximum only).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false
//   ./filter_variants [passes]
// Prints lane-FMAs/s per variant; the register file (two 32-bit banks per lane and cycle plus the
// operand reuse caches) is what separates them.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int DR = 20, TILE_N = 64, TM = 2, THREADS = 128;
constexpr int TILE_FLOATS = (DR + 1) * TILE_N;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long *>(&a);
    rb = *reinterpret_cast<unsigned long long *>(&b);
    rc = *reinterpret_cast<unsigned long long *>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// V0: scalar FFMA, 4 live points per group (round 1)
__device__ __forceinline__ float filter_v0(const float (&a)[TM][DR], const float *T)
{
    float best = -1e30f;
#pragma unroll 1
    for (int g = 0; g < TILE_N / 4; g++) {
        const float *Tg = T + g * 4;
        float acc[TM][4];
        const float4 h = *reinterpret_cast<const float4 *>(Tg + DR * TILE_N);
#pragma unroll
        for (int m = 0; m < TM; m++) { acc[m][0] = h.x; acc[m][1] = h.y; acc[m][2] = h.z; acc[m][3] = h.w; }
#pragma unroll
        for (int k = 0; k < DR; k++) {
            const float4 b = *reinterpret_cast<const float4 *>(Tg + k * TILE_N);
#pragma unroll
            for (int m = 0; m < TM; m++) {
                acc[m][0] = fmaf(a[m][k], b.x, acc[m][0]);
                acc[m][1] = fmaf(a[m][k], b.y, acc[m][1]);
                acc[m][2] = fmaf(a[m][k], b.z, acc[m][2]);
                acc[m][3] = fmaf(a[m][k], b.w, acc[m][3]);
            }
        }
#pragma unroll
        for (int m = 0; m < TM; m++)
            best = fmaxf(best, fmaxf(fmaxf(acc[m][0], acc[m][1]), fmaxf(acc[m][2], acc[m][3])));
    }
    return best;
}

// V1: FFMA2, pairs over live points, proposal coordinate as the scalar operand
__device__ __forceinline__ float filter_v1(const float (&a)[TM][DR], const float *T)
{
    float best = -1e30f;
#pragma unroll 1
    for (int g = 0; g < TILE_N / 4; g++) {
        const float *Tg = T + g * 4;
        float2 acc01[TM], acc23[TM];
        const float4 h = *reinterpret_cast<const float4 *>(Tg + DR * TILE_N);
#pragma unroll
        for (int m = 0; m < TM; m++) { acc01[m] = make_float2(h.x, h.y); acc23[m] = make_float2(h.z, h.w); }
#pragma unroll
        for (int k = 0; k < DR; k++) {
            const float4 b = *reinterpret_cast<const float4 *>(Tg + k * TILE_N);
#pragma unroll
            for (int m = 0; m < TM; m++) {
                const float2 am = make_float2(a[m][k], a[m][k]);
                acc01[m] = ffma2(am, make_float2(b.x, b.y), acc01[m]);
                acc23[m] = ffma2(am, make_float2(b.z, b.w), acc23[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < TM; m++)
            best = fmaxf(best, fmaxf(fmaxf(acc01[m].x, acc01[m].y), fmaxf(acc23[m].x, acc23[m].y)));
    }
    return best;
}

// V2: FFMA2, pairs over the two proposal slots, live coordinate as the scalar operand
__device__ __forceinline__ float filter_v2(const float2 (&a2)[DR], const float *T)
{
    float best = -1e30f;
#pragma unroll 1
    for (int g = 0; g < TILE_N / 4; g++) {
        const float *Tg = T + g * 4;
        float2 acc[4];
        const float4 h = *reinterpret_cast<const float4 *>(Tg + DR * TILE_N);
        acc[0] = make_float2(h.x, h.x); acc[1] = make_float2(h.y, h.y);
        acc[2] = make_float2(h.z, h.z); acc[3] = make_float2(h.w, h.w);
#pragma unroll
        for (int k = 0; k < DR; k++) {
            const float4 b = *reinterpret_cast<const float4 *>(Tg + k * TILE_N);
            acc[0] = ffma2(make_float2(b.x, b.x), a2[k], acc[0]);
            acc[1] = ffma2(make_float2(b.y, b.y), a2[k], acc[1]);
            acc[2] = ffma2(make_float2(b.z, b.z), a2[k], acc[2]);
            acc[3] = ffma2(make_float2(b.w, b.w), a2[k], acc[3]);
        }
        best = fmaxf(best, fmaxf(fmaxf(acc[0].x, acc[1].x), fmaxf(acc[2].x, acc[3].x)));
        best = fmaxf(best, fmaxf(fmaxf(acc[0].y, acc[1].y), fmaxf(acc[2].y, acc[3].y)));
    }
    return best;
}

// V3: FFMA2, pairs over k (even / odd partial sums), all three operands packed; Gray-code walk over
// (slot, live point) so that consecutive instructions share one operand.  Tile layout: k-pairs
// interleaved, T2[kk][n] = (b[2kk][n], b[2kk+1][n]).
__device__ __forceinline__ float filter_v3(const float2 (&a2)[TM][DR / 2], const float *T)
{
    float best = -1e30f;
#pragma unroll 1
    for (int g = 0; g < TILE_N / 4; g++) {
        const float2 *Tg = reinterpret_cast<const float2 *>(T) + g * 4;
        float2 acc[TM][4];
        const float4 h = *reinterpret_cast<const float4 *>(T + DR * TILE_N + g * 4);
#pragma unroll
        for (int m = 0; m < TM; m++) {
            acc[m][0] = make_float2(h.x, 0.f); acc[m][1] = make_float2(h.y, 0.f);
            acc[m][2] = make_float2(h.z, 0.f); acc[m][3] = make_float2(h.w, 0.f);
        }
#pragma unroll
        for (int kk = 0; kk < DR / 2; kk++) {
            const float4 b01 = *reinterpret_cast<const float4 *>(Tg + kk * TILE_N);
            const float4 b23 = *reinterpret_cast<const float4 *>(Tg + kk * TILE_N + 2);
            const float2 b0 = make_float2(b01.x, b01.y), b1 = make_float2(b01.z, b01.w);
            const float2 b2 = make_float2(b23.x, b23.y), b3 = make_float2(b23.z, b23.w);
            acc[0][0] = ffma2(a2[0][kk], b0, acc[0][0]);
            acc[0][1] = ffma2(a2[0][kk], b1, acc[0][1]);
            acc[1][1] = ffma2(a2[1][kk], b1, acc[1][1]);
            acc[1][2] = ffma2(a2[1][kk], b2, acc[1][2]);
            acc[0][2] = ffma2(a2[0][kk], b2, acc[0][2]);
            acc[0][3] = ffma2(a2[0][kk], b3, acc[0][3]);
            acc[1][3] = ffma2(a2[1][kk], b3, acc[1][3]);
            acc[1][0] = ffma2(a2[1][kk], b0, acc[1][0]);
        }
#pragma unroll
        for (int m = 0; m < TM; m++) {
            const float s0 = acc[m][0].x + acc[m][0].y, s1 = acc[m][1].x + acc[m][1].y;
            const float s2 = acc[m][2].x + acc[m][2].y, s3 = acc[m][3].x + acc[m][3].y;
            best = fmaxf(best, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
        }
    }
    return best;
}

// V4: like V2 (pairs over slots) but 8 live points per group: two LDS.128 per k, 8 accumulator pairs
__device__ __forceinline__ float filter_v4(const float2 (&a2)[DR], const float *T)
{
    float best = -1e30f;
#pragma unroll 1
    for (int g = 0; g < TILE_N / 8; g++) {
        const float *Tg = T + g * 8;
        float2 acc[8];
        const float4 h0 = *reinterpret_cast<const float4 *>(Tg + DR * TILE_N);
        const float4 h1 = *reinterpret_cast<const float4 *>(Tg + DR * TILE_N + 4);
        acc[0] = make_float2(h0.x, h0.x); acc[1] = make_float2(h0.y, h0.y);
        acc[2] = make_float2(h0.z, h0.z); acc[3] = make_float2(h0.w, h0.w);
        acc[4] = make_float2(h1.x, h1.x); acc[5] = make_float2(h1.y, h1.y);
        acc[6] = make_float2(h1.z, h1.z); acc[7] = make_float2(h1.w, h1.w);
#pragma unroll
        for (int k = 0; k < DR; k++) {
            const float4 b = *reinterpret_cast<const float4 *>(Tg + k * TILE_N);
            const float4 c = *reinterpret_cast<const float4 *>(Tg + k * TILE_N + 4);
            acc[0] = ffma2(make_float2(b.x, b.x), a2[k], acc[0]);
            acc[1] = ffma2(make_float2(b.y, b.y), a2[k], acc[1]);
            acc[2] = ffma2(make_float2(b.z, b.z), a2[k], acc[2]);
            acc[3] = ffma2(make_float2(b.w, b.w), a2[k], acc[3]);
            acc[4] = ffma2(make_float2(c.x, c.x), a2[k], acc[4]);
            acc[5] = ffma2(make_float2(c.y, c.y), a2[k], acc[5]);
            acc[6] = ffma2(make_float2(c.z, c.z), a2[k], acc[6]);
            acc[7] = ffma2(make_float2(c.w, c.w), a2[k], acc[7]);
        }
        float bx = fmaxf(fmaxf(acc[0].x, acc[1].x), fmaxf(acc[2].x, acc[3].x));
        float by = fmaxf(fmaxf(acc[0].y, acc[1].y), fmaxf(acc[2].y, acc[3].y));
        bx = fmaxf(bx, fmaxf(fmaxf(acc[4].x, acc[5].x), fmaxf(acc[6].x, acc[7].x)));
        by = fmaxf(by, fmaxf(fmaxf(acc[4].y, acc[5].y), fmaxf(acc[6].y, acc[7].y)));
        best = fmaxf(best, fmaxf(bx, by));
    }
    return best;
}

template <int V>
__global__ void __launch_bounds__(THREADS, 4) k_variant(const float *tiles, const float *cand, int passes, float *out)
{
    __shared__ __align__(16) float tbuf[2][TILE_FLOATS];
    for (int i = threadIdx.x; i < 2 * TILE_FLOATS; i += THREADS) (&tbuf[0][0])[i] = tiles[i];
    const int t = blockIdx.x * THREADS + threadIdx.x;
    float a[TM][DR];
#pragma unroll
    for (int m = 0; m < TM; m++)
#pragma unroll
        for (int k = 0; k < DR; k++) a[m][k] = cand[(size_t)(t * TM + m) * DR + k];
    float2 a2[DR];
#pragma unroll
    for (int k = 0; k < DR; k++) a2[k] = make_float2(a[0][k], a[1][k]);
    float2 ak[TM][DR / 2];
#pragma unroll
    for (int m = 0; m < TM; m++)
#pragma unroll
        for (int kk = 0; kk < DR / 2; kk++) ak[m][kk] = make_float2(a[m][2 * kk], a[m][2 * kk + 1]);
    __syncthreads();
    float best = -1e30f;
    for (int p = 0; p < passes; p++) {
        const float *T = tbuf[p & 1];
        float r;
        if (V == 0) r = filter_v0(a, T);
        else if (V == 1) r = filter_v1(a, T);
        else if (V == 2) r = filter_v2(a2, T);
        else if (V == 3) r = filter_v3(ak, T);
        else r = filter_v4(a2, T);
        best = fmaxf(best, r);
        __syncthreads();   // the real kernel hands tiles over at a block barrier too
    }
    if (best == 12345.678f) out[t] = best;
}

template <int V>
double run(const float *tiles, const float *cand, int passes, float *out, int blocks)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_variant<V><<<blocks, THREADS>>>(tiles, cand, passes, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)blocks * THREADS * TM * passes * TILE_N * DR / (ms * 1e-3);
        if (rep && rate > best) best = rate;
    }
    return best;
}

int main(int argc, char **argv)
{
    const int passes = argc > 1 ? atoi(argv[1]) : 2000;
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4;
    float *tiles, *cand, *out;
    cudaMalloc(&tiles, 2 * TILE_FLOATS * sizeof(float));
    cudaMalloc(&cand, (size_t)blocks * THREADS * TM * DR * sizeof(float));
    cudaMalloc(&out, (size_t)blocks * THREADS * sizeof(float));
    cudaMemset(tiles, 0, 2 * TILE_FLOATS * sizeof(float));
    cudaMemset(cand, 0, (size_t)blocks * THREADS * TM * DR * sizeof(float));
    printf("{\"passes\": %d, \"blocks\": %d", passes, blocks);
    printf(", \"v0_scalar_ffma\": %.4g", run<0>(tiles, cand, passes, out, blocks));
    printf(", \"v1_ffma2_pair_points\": %.4g", run<1>(tiles, cand, passes, out, blocks));
    printf(", \"v2_ffma2_pair_slots\": %.4g", run<2>(tiles, cand, passes, out, blocks));
    printf(", \"v3_ffma2_pair_k_gray\": %.4g", run<3>(tiles, cand, passes, out, blocks));
    printf(", \"v4_ffma2_pair_slots_8pts\": %.4g", run<4>(tiles, cand, passes, out, blocks));
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
