// probe_store.cu -- what write bandwidth does B200 HBM3e sustain for the OUTPUT PATTERNS of the FK kernel, with no
// arithmetic at all?  348 doubles per configuration (25 links x 12 + 6 x 8 Jacobian), N configurations.
//   fill      contiguous grid-stride stores (the roof)
//   soa       x[comp * N + n]: a CTA tile of 128 configurations writes 1 KB to each of 348 rows (plain / st.cs)
//   tiled     x[((n / 32) * 348 + comp) * 32 + n % 32]: each warp writes one contiguous 89 KB block
//   soa_bulk  as soa, staged per warp through shared memory, 256-B cp.async.bulk per component row
//   tiled_bulk as tiled, staged per warp: 12 components x 32 lanes = 3 KB per cp.async.bulk
//   soa_cta_bulk  CTA-wide staging [12][128], one 1-KB cp.async.bulk per component row
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/_bin/probe_store profiles/probe_store.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int REC = 348, BS = 128;

__global__ void k_fill(double *x, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) x[i] = 1.0;
}
template <bool CS, bool TILED>
__global__ void __launch_bounds__(BS) k_pattern(double *x, long long N) {
    const long long n_tiles = N / BS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n = tile * BS + threadIdx.x;
        double *p = TILED ? x + (n >> 5) * (long long)(REC * 32) + (n & 31) : x + n;
        const size_t es = TILED ? 32 : (size_t)N;
        const double v = (double)n;
        #pragma unroll 12
        for (int c = 0; c < REC; ++c) {
            if (CS) __stcs(p + c * es, v + c); else p[c * es] = v + c;
        }
    }
}
// the same pattern with the things the real kernel has and the plain probe lacks, one at a time:
//   READQ    8 coalesced loads per configuration before the stores (mixed read / write traffic)
//   DELAY    a dependent DFMA chain of that length in front of every group of 12 stores (the stores of a warp are
//            spread over the computation instead of being issued back to back)
//   PERMUTE  the 29 groups of 12 components are visited in a scattered order (DFS order != output order)
template <bool TILED, bool READQ, int DELAY, bool PERMUTE>
__global__ void __launch_bounds__(BS) k_real(double *x, const double *q, long long N) {
    const long long n_tiles = N / BS;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n = tile * BS + threadIdx.x;
        double *p = TILED ? x + (n >> 5) * (long long)(REC * 32) + (n & 31) : x + n;
        const size_t es = TILED ? 32 : (size_t)N;
        double v = (double)n;
        if (READQ) {
            const double *qp = TILED ? q + (n >> 5) * (long long)(8 * 32) + (n & 31) : q + n;
            #pragma unroll
            for (int c = 0; c < 8; ++c) v += qp[c * es];
        }
        #pragma unroll 1
        for (int g0 = 0; g0 < 29; ++g0) {
            const int g = PERMUTE ? (g0 * 11 + 7) % 29 : g0;
            #pragma unroll
            for (int d = 0; d < DELAY; ++d) v = fma(v, 1.0000001, 1e-9);
            #pragma unroll
            for (int c = 0; c < 12; ++c) p[(g * 12 + c) * es] = v + c;
        }
    }
}

// READQ with the reads clustered in time: every PF-th tile each CTA prefetches the configurations of its tiles
// k + PF .. k + 2 PF - 1 into L2 (evict_last); the CTAs run in near lock-step, so the DRAM sees the reads of a whole
// period as one burst instead of a trickle that keeps interrupting the write drain.
template <bool TILED, int PF>
__global__ void __launch_bounds__(BS) k_real_pf(double *x, const double *q, long long N) {
    const long long n_tiles = N / BS;
    long long k = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
        if (k % PF == 0) {
            // lines of 128 B: per tile 8 rows x 8 lines (SoA) or one contiguous 8 KB block (tiled: 64 lines)
            for (int i = threadIdx.x; i < PF * 64; i += BS) {
                const long long t2 = tile + (long long)(PF + i / 64) * gridDim.x;
                if (t2 < n_tiles) {
                    const int l = i % 64;
                    const double *a = TILED ? q + t2 * (long long)(BS * 8) + l * 16 : q + (long long)(l / 8) * N + t2 * BS + (l % 8) * 16;
                    asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
                }
            }
        }
        const long long n = tile * BS + threadIdx.x;
        double *p = TILED ? x + (n >> 5) * (long long)(REC * 32) + (n & 31) : x + n;
        const size_t es = TILED ? 32 : (size_t)N;
        double v = (double)n;
        const double *qp = TILED ? q + (n >> 5) * (long long)(8 * 32) + (n & 31) : q + n;
        #pragma unroll
        for (int c = 0; c < 8; ++c) v += qp[c * es];
        #pragma unroll 1
        for (int g = 0; g < 29; ++g) {
            #pragma unroll
            for (int c = 0; c < 12; ++c) p[(g * 12 + c) * es] = v + c;
        }
    }
}

// More ways of getting the 8 input doubles per configuration in (SoA outputs throughout):
//   MODE 0  per-thread loads with a cache operator (CACHE: 0 plain, 1 .cg, 2 .cs, 3 .nc)
//   MODE 1  warp 0 loads the whole tile (8 rows x 1 KB) with 16-byte loads into shared memory, then a barrier
//   MODE 2  every RT-th tile the CTA loads the inputs of RT tiles into shared memory (8 rows x RT KB)
//   MODE 3  the inputs are stored tiled (one contiguous 8 KB block per tile of 128), outputs stay SoA
template <int MODE, int CACHE, int RT>
__global__ void __launch_bounds__(BS) k_real2(double *x, const double *q, long long N) {
    __shared__ __align__(16) double sq[(MODE == 2 ? RT : 1) * 8 * BS];
    const long long n_tiles = N / BS;
    long long k = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
        const long long n = tile * BS + threadIdx.x;
        double v = (double)n;
        if (MODE == 0) {
            #pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double *a = q + (long long)c * N + n;
                double t;
                if (CACHE == 1) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(t) : "l"(a));
                else if (CACHE == 2) asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(t) : "l"(a));
                else if (CACHE == 3) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(t) : "l"(a));
                else t = *a;
                v += t;
            }
        } else if (MODE == 1) {
            __syncthreads();
            if (threadIdx.x < 32)
                #pragma unroll
                for (int i = 0; i < 16; ++i) {        // 8 rows x 64 double2 = 512 double2, 32 lanes x 16
                    const int e = i * 32 + threadIdx.x, row = e / 64, col = e % 64;
                    reinterpret_cast<double2 *>(sq)[row * 64 + col] = reinterpret_cast<const double2 *>(q + (long long)row * N + tile * BS)[col];
                }
            __syncthreads();
            #pragma unroll
            for (int c = 0; c < 8; ++c) v += sq[c * BS + threadIdx.x];
        } else if (MODE == 2) {
            if (k % RT == 0) {
                __syncthreads();
                for (int r = 0; r < RT; ++r) {
                    const long long t2 = tile + (long long)r * gridDim.x;
                    if (t2 < n_tiles)
                        #pragma unroll
                        for (int c = 0; c < 8; ++c) sq[(r * 8 + c) * BS + threadIdx.x] = q[(long long)c * N + t2 * BS + threadIdx.x];
                }
                __syncthreads();
            }
            #pragma unroll
            for (int c = 0; c < 8; ++c) v += sq[((k % RT) * 8 + c) * BS + threadIdx.x];
        } else {
            const double *qp = q + tile * (long long)(BS * 8) + threadIdx.x;
            #pragma unroll
            for (int c = 0; c < 8; ++c) v += qp[c * BS];
        }
        double *p = x + n;
        #pragma unroll 1
        for (int g = 0; g < 29; ++g) {
            #pragma unroll
            for (int c = 0; c < 12; ++c) p[(size_t)(g * 12 + c) * N] = v + c;
        }
    }
}

// Inputs of RT tiles per batch, double-buffered in DYNAMIC shared memory with cp.async (the next batch is in flight
// while the current one is processed); optionally every CTA waits at a grid-wide barrier before issuing the loads of
// a batch, so that all SMs read at the same moment (one read burst per batch for the whole GPU).
__device__ unsigned g_bar_count = 0;
__device__ volatile unsigned g_bar_gen = 0;
__device__ __forceinline__ void grid_barrier(unsigned n_cta) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned gen = g_bar_gen;
        __threadfence();
        if (atomicAdd(&g_bar_count, 1) == n_cta - 1) { g_bar_count = 0; __threadfence(); g_bar_gen = gen + 1; }
        else while (g_bar_gen == gen) __nanosleep(100);
    }
    __syncthreads();
}
template <bool GRIDSYNC>
__global__ void __launch_bounds__(BS) k_batched(double *x, const double *q, long long N, int RT) {
    extern __shared__ __align__(16) double sqd[];          // [2][RT][8][BS]
    const long long n_tiles = N / BS;
    const long long my_tiles = (long long)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_batches = (my_tiles + RT - 1) / RT;
    const long long max_tiles = (n_tiles + gridDim.x - 1) / gridDim.x, max_batches = (max_tiles + RT - 1) / RT;
    auto issue = [&](long long b, int buf) {
        for (int r = 0; r < RT; ++r) {
            const long long k = b * RT + r;
            if (k < my_tiles) {
                const long long t2 = blockIdx.x + k * gridDim.x;
                #pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const double *src = q + (long long)c * N + t2 * BS + threadIdx.x;
                    double *dst = sqd + ((size_t)(buf * RT + r) * 8 + c) * BS + threadIdx.x;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
                }
            }
        }
        asm volatile("cp.async.commit_group;");
    };
    if (GRIDSYNC) grid_barrier(gridDim.x);
    issue(0, 0);
    for (long long b = 0; b < max_batches; ++b) {
        const int buf = (int)(b & 1);
        if (GRIDSYNC) grid_barrier(gridDim.x);
        issue(b + 1, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (b < n_batches)
            for (int r = 0; r < RT; ++r) {
                const long long k = b * RT + r;
                if (k >= my_tiles) break;
                const long long n = (blockIdx.x + k * gridDim.x) * BS + threadIdx.x;
                double v = (double)n;
                #pragma unroll
                for (int c = 0; c < 8; ++c) v += sqd[((size_t)(buf * RT + r) * 8 + c) * BS + threadIdx.x];
                double *p = x + n;
                #pragma unroll 1
                for (int g = 0; g < 29; ++g) {
                    #pragma unroll
                    for (int c = 0; c < 12; ++c) p[(size_t)(g * 12 + c) * N] = v + c;
                }
            }
        __syncthreads();
    }
}

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_) : "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// per-warp staging: CH components x 32 lanes, double buffered
template <bool TILED, int CH>
__global__ void __launch_bounds__(BS) k_warp_bulk(double *x, long long N) {
    __shared__ __align__(128) double stage[BS / 32][2][CH][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_tiles = N / BS;
    int buf = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n0 = tile * BS + w * 32;          // first configuration of this warp
        const double v = (double)(n0 + lane);
        for (int c0 = 0; c0 < REC; c0 += CH) {
            if (lane == 0) bulk_wait_read<1>();            // the buffer we are about to overwrite has been read
            __syncwarp();
            const int ch = min(CH, REC - c0);
            #pragma unroll
            for (int c = 0; c < CH; ++c) if (c < ch) stage[w][buf][c][lane] = v + c0 + c;
            fence_async();
            __syncwarp();
            if (lane == 0) {
                if (TILED) bulk_store(x + (n0 >> 5) * (long long)(REC * 32) + (long long)c0 * 32, &stage[w][buf][0][0], ch * 32 * 8);
                else
                    for (int c = 0; c < ch; ++c) bulk_store(x + (long long)(c0 + c) * N + n0, &stage[w][buf][c][0], 32 * 8);
                bulk_commit();
            }
            buf ^= 1;
        }
    }
    if (lane == 0) bulk_wait_read<0>();
}
// CTA-wide staging [CH][128]: one 1-KB bulk store per component row, issued by warp 0's lanes
template <int CH>
__global__ void __launch_bounds__(BS) k_cta_bulk(double *x, long long N) {
    __shared__ __align__(128) double stage[2][CH][BS];
    const long long n_tiles = N / BS;
    int buf = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long n0 = tile * BS;
        const double v = (double)(n0 + threadIdx.x);
        for (int c0 = 0; c0 < REC; c0 += CH) {
            if (threadIdx.x < CH) bulk_wait_read<1>();
            __syncthreads();
            #pragma unroll
            for (int c = 0; c < CH; ++c) stage[buf][c][threadIdx.x] = v + c0 + c;
            fence_async();
            __syncthreads();
            if (threadIdx.x < CH) {
                bulk_store(x + (long long)(c0 + threadIdx.x) * N + n0, &stage[buf][threadIdx.x][0], BS * 8);
                bulk_commit();
            }
            buf ^= 1;
        }
    }
    if (threadIdx.x < CH) bulk_wait_read<0>();
}

template <typename F> float timeit(F launch, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main(int argc, char **argv) {
    const long long N = argc > 1 ? atoll(argv[1]) : (1ll << 22);
    const long long total = N * REC;
    double *x;
    CK(cudaMalloc(&x, total * 8));
    int n_sm; CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
    const double gb = total * 8 / 1e9;
    auto rep = [&](const char *name, float ms) { printf("%-34s %8.3f ms  %8.1f GB/s\n", name, ms, gb / (ms * 1e-3)); };
    rep("fill (contiguous)", timeit([&] { k_fill<<<n_sm * 16, 256>>>(x, total); }, 5));
    for (int occ : {3, 4, 8, 16}) {
        char nm[64];
        snprintf(nm, sizeof nm, "soa plain        grid %d/SM", occ); rep(nm, timeit([&] { k_pattern<false, false><<<n_sm * occ, BS>>>(x, N); }, 5));
        snprintf(nm, sizeof nm, "soa st.cs        grid %d/SM", occ); rep(nm, timeit([&] { k_pattern<true, false><<<n_sm * occ, BS>>>(x, N); }, 5));
        snprintf(nm, sizeof nm, "tiled plain      grid %d/SM", occ); rep(nm, timeit([&] { k_pattern<false, true><<<n_sm * occ, BS>>>(x, N); }, 5));
        snprintf(nm, sizeof nm, "tiled st.cs      grid %d/SM", occ); rep(nm, timeit([&] { k_pattern<true, true><<<n_sm * occ, BS>>>(x, N); }, 5));
    }
    {
        double *q;
        CK(cudaMalloc(&q, N * 8 * 8));
        CK(cudaMemset(q, 0, N * 8 * 8));
        const int g3 = n_sm * 3;
        rep("soa   real: base (3/SM)", timeit([&] { k_real<false, false, 0, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +read q", timeit([&] { k_real<false, true, 0, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +delay 16", timeit([&] { k_real<false, false, 16, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +delay 64", timeit([&] { k_real<false, false, 64, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +permute", timeit([&] { k_real<false, false, 0, true><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: all (delay 16)", timeit([&] { k_real<false, true, 16, true><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: ld.cg", timeit([&] { k_real2<0, 1, 1><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: ld.cs", timeit([&] { k_real2<0, 2, 1><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: ld.nc no_allocate", timeit([&] { k_real2<0, 3, 1><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: warp 0 loads tile -> smem", timeit([&] { k_real2<1, 0, 1><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: 4 tiles at a time -> smem", timeit([&] { k_real2<2, 0, 4><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: 5 tiles at a time -> smem", timeit([&] { k_real2<2, 0, 5><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: tiled q, soa out", timeit([&] { k_real2<3, 0, 1><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real2: plain, 1 CTA/SM", timeit([&] { k_real2<0, 0, 1><<<n_sm, BS>>>(x, q, N); }, 5));
        rep("soa   real2: plain, 2 CTA/SM", timeit([&] { k_real2<0, 0, 1><<<n_sm * 2, BS>>>(x, q, N); }, 5));
        rep("soa   real2: plain, 6 CTA/SM", timeit([&] { k_real2<0, 0, 1><<<n_sm * 6, BS>>>(x, q, N); }, 5));
        for (int occ : {1, 2}) for (int RT : {2, 4, 6, 12}) {
            const size_t smem = (size_t)2 * RT * 8 * BS * 8;
            if (smem * occ > 220 * 1024) continue;
            CK(cudaFuncSetAttribute(k_batched<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(k_batched<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            char nm[80];
            snprintf(nm, sizeof nm, "soa batched cp.async RT=%d, %d CTA/SM", RT, occ);
            rep(nm, timeit([&] { k_batched<false><<<n_sm * occ, BS, smem>>>(x, q, N, RT); }, 5));
            snprintf(nm, sizeof nm, "soa batched+gridsync RT=%d, %d CTA/SM", RT, occ);
            rep(nm, timeit([&] { k_batched<true><<<n_sm * occ, BS, smem>>>(x, q, N, RT); }, 5));
        }
        rep("soa   real: +read q, L2 prefetch P=2", timeit([&] { k_real_pf<false, 2><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +read q, L2 prefetch P=4", timeit([&] { k_real_pf<false, 4><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +read q, L2 prefetch P=8", timeit([&] { k_real_pf<false, 8><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +read q, L2 prefetch P=16", timeit([&] { k_real_pf<false, 16><<<g3, BS>>>(x, q, N); }, 5));
        rep("soa   real: +read q, L2 prefetch P=32", timeit([&] { k_real_pf<false, 32><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: +read q, L2 prefetch P=4", timeit([&] { k_real_pf<true, 4><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: +read q, L2 prefetch P=16", timeit([&] { k_real_pf<true, 16><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: base (3/SM)", timeit([&] { k_real<true, false, 0, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: +read q", timeit([&] { k_real<true, true, 0, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: +delay 16", timeit([&] { k_real<true, false, 16, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: +delay 64", timeit([&] { k_real<true, false, 64, false><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: +permute", timeit([&] { k_real<true, false, 0, true><<<g3, BS>>>(x, q, N); }, 5));
        rep("tiled real: all (delay 16)", timeit([&] { k_real<true, true, 16, true><<<g3, BS>>>(x, q, N); }, 5));
        CK(cudaFree(q));
    }
    for (int occ : {3, 8}) {
        char nm[64];
        snprintf(nm, sizeof nm, "soa warp-bulk 256B  grid %d/SM", occ); rep(nm, timeit([&] { k_warp_bulk<false, 12><<<n_sm * occ, BS>>>(x, N); }, 5));
        snprintf(nm, sizeof nm, "tiled warp-bulk 3KB grid %d/SM", occ); rep(nm, timeit([&] { k_warp_bulk<true, 12><<<n_sm * occ, BS>>>(x, N); }, 5));
        snprintf(nm, sizeof nm, "tiled warp-bulk 6KB grid %d/SM", occ); rep(nm, timeit([&] { k_warp_bulk<true, 24><<<n_sm * occ, BS>>>(x, N); }, 5));
        snprintf(nm, sizeof nm, "soa cta-bulk 1KB    grid %d/SM", occ); rep(nm, timeit([&] { k_cta_bulk<12><<<n_sm * occ, BS>>>(x, N); }, 5));
    }
    CK(cudaFree(x));
    return 0;
}
