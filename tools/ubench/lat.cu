// Latency microbenchmarks for the mixing walker's design (one warp, dependent chains; cycles per step).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

__global__ void k_chain(int which, unsigned seed, long long *out, unsigned *sink) {
    __shared__ unsigned sm[4096];
    __shared__ volatile unsigned flag[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = threadIdx.x; q < 4096; q += blockDim.x) sm[q] = (q * 97u + 13u) & 4095u;
    if (threadIdx.x < 64) flag[threadIdx.x] = 0;
    __syncthreads();
    unsigned x = seed + lane;
    float f = (float)lane;
    long long t0 = 0, t1 = 0;
    if (which == 0) {          // LOP3/IADD chain
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = (x ^ (x << 1)) + 0x9e3779b9u;   // 2-3 dependent ALU ops
        t1 = clk();
    } else if (which == 1) {   // predicate -> VOTE -> register
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = __ballot_sync(0xffffffffu, (x >> lane) & 1u) + n;
        t1 = clk();
    } else if (which == 2) {   // REDUX.OR
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = __reduce_or_sync(0xffffffffu, x & (1u << lane)) + n + lane;
        t1 = clk();
    } else if (which == 3) {   // SHFL
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = __shfl_sync(0xffffffffu, x, (x + n) & 31) + 1;
        t1 = clk();
    } else if (which == 4) {   // LDS pointer chase
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = sm[x & 4095u];
        t1 = clk();
    } else if (which == 5) {   // STS then dependent LDS of the same word
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) { sm[(lane * 33 + n) & 4095] = x; x = sm[(lane * 33 + n) & 4095] + 1; }
        t1 = clk();
    } else if (which == 6) {   // FADD chain
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) f = f + 1.0f;
        t1 = clk();
        x = __float_as_uint(f);
    } else if (which == 7) {   // POPC chain
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = __popc(x) + n;
        t1 = clk();
    } else if (which == 8) {   // LDS.U16 + LDS.U8 + PRMT chain (image decode)
        const unsigned short *s16 = (const unsigned short *)sm; const unsigned char *s8 = (const unsigned char *)sm;
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = __byte_perm((unsigned)s16[x & 8191u], (unsigned)s8[x & 16383u], 0x1045) >> 7;
        t1 = clk();
    } else if (which == 9) {   // FSETP -> VOTE
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) { unsigned b = __ballot_sync(0xffffffffu, f > (float)n); f = f + (float)(b & 1u); }
        t1 = clk();
        x = __float_as_uint(f);
    } else if (which == 10) {  // __syncwarp chain with shared store/load across lanes
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1; __syncwarp(); }
        t1 = clk();
    } else if (which == 11) {  // two-warp ping-pong through volatile shared flags (warps 0 and 1: different SMSPs)
        if (warp < 2) {
            t0 = clk();
            for (int n = 1; n <= N / 4; n++) {
                if (warp == 0) { if (lane == 0) { flag[0] = n; while (flag[32] != (unsigned)n) ; } }
                else { if (lane == 0) { while (flag[0] != (unsigned)n) ; flag[32] = n; } }
                __syncwarp();
            }
            t1 = clk();
        }
    } else if (which == 12) {  // bar.sync with all warps of the CTA
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N / 4; n++) { __syncthreads(); x += n; }
        t1 = clk();
    } else if (which == 13) {  // named barrier, 2 warps (0 and 1)
        if (warp < 2) {
            t0 = clk();
            for (int n = 0; n < N / 4; n++) { asm volatile("bar.sync 1, 64;" ::: "memory"); x += n; }
            t1 = clk();
        }
    } else if (which == 14) {  // ping-pong with st.release / ld.acquire (cta scope)
        if (warp < 2) {
            unsigned a0 = (unsigned)__cvta_generic_to_shared((const void *)&flag[0]), a1 = (unsigned)__cvta_generic_to_shared((const void *)&flag[32]);
            t0 = clk();
            for (int n = 1; n <= N / 4; n++) {
                if (lane == 0) {
                    unsigned v;
                    if (warp == 0) { asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(a0), "r"(n) : "memory");
                        do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a1) : "memory"); } while (v != (unsigned)n); }
                    else { do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a0) : "memory"); } while (v != (unsigned)n);
                        asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(a1), "r"(n) : "memory"); }
                }
                __syncwarp();
            }
            t1 = clk();
        }
    } else if (which == 15) {  // LDS.64 pointer chase
        const uint2 *s2 = (const uint2 *)sm;
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) { uint2 v = s2[x & 2047u]; x = v.x ^ (v.y & 1u); }
        t1 = clk();
    } else if (which == 16) {  // funnel-shift rotate + lop chain
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = __funnelshift_r(x, x, n & 31) ^ 0x5bd1e995u;
        t1 = clk();
    } else if (which == 17) {  // ISETP -> SEL chain (predicate to ALU)
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) x = (x & 0x10u) ? (x >> 1) + n : (x << 1) ^ n;
        t1 = clk();
    } else if (which == 18) {  // two independent VOTE chains interleaved (throughput)
        unsigned y = seed * 3 + lane;
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N; n++) { x = __ballot_sync(0xffffffffu, (x >> lane) & 1u) + n; y = __ballot_sync(0xffffffffu, (y >> lane) & 1u) ^ n; }
        t1 = clk();
        x ^= y;
    } else if (which == 19) {  // match_any
        t0 = clk();
#pragma unroll 16
        for (int n = 0; n < N / 4; n++) x = __match_any_sync(0xffffffffu, x & 7u) + n;
        t1 = clk();
    }
    if (lane == 0 && warp == 0) out[which] = t1 - t0;
    sink[threadIdx.x] = x;
}

int main() {
    long long *out; unsigned *sink;
    cudaMallocManaged(&out, 64 * sizeof(long long)); cudaMalloc(&sink, 1024 * 4);
    const char *names[] = {"alu chain (3 ops)", "pred->VOTE->reg", "REDUX.OR", "SHFL", "LDS chase", "STS->LDS same addr", "FADD", "POPC",
                           "LDS.U16+U8+PRMT", "FSETP->VOTE->FADD", "STS syncwarp LDS syncwarp", "2-warp pingpong volatile (round trip)",
                           "bar.sync 128 thr", "bar.sync named 2 warps", "2-warp pingpong acq/rel (round trip)", "LDS.64 chase", "SHF rot+LOP",
                           "ISETP->SEL chain", "2 interleaved VOTE chains", "match_any"};
    const int steps[] = {N, N, N, N, N, N, N, N, N, N, N, N / 4, N / 4, N / 4, N / 4, N, N, N, N, N / 4};
    for (int rep = 0; rep < 2; rep++)
        for (int w = 0; w < 20; w++) {
            k_chain<<<1, 128>>>(w, 12345u, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %d %s\n", w, cudaGetErrorString(e)); return 1; }
            if (rep == 1) printf("%-42s %8.1f cycles/step\n", names[w], (double)out[w] / steps[w]);
        }
    return 0;
}
