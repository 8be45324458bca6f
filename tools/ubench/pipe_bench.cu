// pipe_bench.cu -- instruction-mix throughput probes for the K1 inner loops (development aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITER 4096
#define ILP 8

template<int V>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, uint32_t c65536, uint32_t m1, uint32_t one)
{
    uint32_t x[ILP];
#pragma unroll
    for(int i = 0; i < ILP; ++i) x[i] = seed + threadIdx.x * 977u + i * 131u + blockIdx.x;
    uint32_t acc = 0;
    for(int it = 0; it < ITER; ++it) {
#pragma unroll
        for(int i = 0; i < ILP; ++i) {
            uint32_t v = x[i];
            if(V == 0) {            // hash round: SHF + LOP3 + IMAD  (2 ALU + 1 FMA)
                v = ((v >> 16) ^ v) * 0x45d9f3bu;
                v = ((v >> 16) ^ v) * 0x45d9f3bu;
            } else if(V == 1) {     // hash round: IMAD.HI + LOP3 + IMAD (1 ALU + 2 FMA)
                uint32_t h;
                asm("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(v), "r"(c65536));
                v = (h ^ v) * 0x45d9f3bu;
                asm("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(v), "r"(c65536));
                v = (h ^ v) * 0x45d9f3bu;
            } else if(V == 2) {     // compare-exchange: 2 x VIMNMX
                const uint32_t o = x[(i + 1) % ILP];
                const uint32_t lo = min(v, o), hi = max(v, o);
                v = lo; x[(i + 1) % ILP] = hi;
            } else if(V == 3) {     // compare-exchange: VIMNMX + 2 x IMAD
                const uint32_t o = x[(i + 1) % ILP];
                const uint32_t lo = min(v, o);
                uint32_t s, hi;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s) : "r"(v), "r"(one), "r"(o));
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(lo), "r"(m1), "r"(s));
                v = lo; x[(i + 1) % ILP] = hi;
            } else if(V == 4) {     // funnel shift pair (2 ALU)
                v = __funnelshift_l(v, x[(i + 1) % ILP], 6) ^ __funnelshift_r(v, x[(i + 3) % ILP], 10);
            } else if(V == 5) {     // pure IMAD chain x2 (FMA pipe rate)
                v = v * 0x45d9f3bu + 12345u;
                v = v * 0x45d9f3bu + 12345u;
            } else if(V == 6) {     // pure IMAD.HI x2
                uint32_t h;
                asm("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(v), "r"(c65536 + 12345u));
                asm("mul.hi.u32 %0, %1, %2;" : "=r"(v) : "r"(h + 0u), "r"(m1));
                v += h;
            } else if(V == 7) {     // pure LOP3 x2 (ALU rate)
                v = (v ^ x[(i + 1) % ILP]) & (x[(i + 2) % ILP] | 0x55u);
                v = (v | x[(i + 3) % ILP]) ^ (x[(i + 5) % ILP] & 0x33u);
            } else if(V == 8) {     // mixed: 1 LOP3 + 1 IMAD
                v = (v ^ x[(i + 1) % ILP]) * 0x45d9f3bu + x[(i + 2) % ILP];
            } else if(V >= 10 && V <= 13) {   // 6 independent-ish LOP3 + FMA-pipe candidates
                uint32_t a = x[(i + 1) % ILP], b = x[(i + 2) % ILP], c = x[(i + 3) % ILP];
                uint32_t l = (v ^ a) & (b | 0x55u);
                l = (l | c) ^ (a & 0x33u);
                l = (l ^ b) | (c & 0x0Fu);
                uint32_t h = 0;
                if(V == 10) {
                    uint32_t h1, h2;
                    asm("mul.hi.u32 %0, %1, %2;" : "=r"(h1) : "r"(a), "r"(c65536));
                    asm("mul.hi.u32 %0, %1, %2;" : "=r"(h2) : "r"(b), "r"(c65536));
                    h = h1 + h2;
                } else if(V == 12) {
                    unsigned long long d1, d2;
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d1) : "r"(a), "r"(one), "l"((unsigned long long)c << 7));
                    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d2) : "r"(b), "r"(one), "l"((unsigned long long)a << 9));
                    h = uint32_t(d1 >> 32) + uint32_t(d2 >> 32);
                } else if(V == 13) {
                    uint32_t h1, h2, h3, h4;
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h1) : "r"(a), "r"(c65536), "r"(b));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h2) : "r"(b), "r"(m1), "r"(c));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h3) : "r"(c), "r"(c65536), "r"(a));
                    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h4) : "r"(h1), "r"(m1), "r"(h2));
                    h = h3 + h4;
                }
                v = l + h;
            } else if(V == 9) {     // funnel via IMAD.HI + IMAD: f = A0*4^r + hi(A1*4^r)
                uint32_t h;
                asm("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(x[(i + 1) % ILP]), "r"(c65536));
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(v), "r"(c65536), "r"(h));
            }
            x[i] = v;
        }
    }
#pragma unroll
    for(int i = 0; i < ILP; ++i) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template<int V>
void run(const char* name, double ops_per_inner, uint32_t* d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8;
    probe<V><<<grid, 256>>>(d, 1u, 65536u, 0xFFFFFFFFu, 1u);
    cudaEventRecord(e0);
    probe<V><<<grid, 256>>>(d, 2u, 65536u, 0xFFFFFFFFu, 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inner = double(grid) * 8 * ITER * ILP;      // warp-level inner bodies
    const double cyc = ms * 1e-3 * 1.92e9 * 148 * 4;               // SMSP cycles available (approx clock)
    printf("%-44s %8.3f ms  %6.2f SMSP-cycles per inner body (%.0f listed ops)\n", name, ms, cyc / warp_inner, ops_per_inner);
}

int main()
{
    uint32_t* d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("hash 2 rounds: SHF+LOP3+IMAD", 6, d);
    run<1>("hash 2 rounds: IMAD.HI+LOP3+IMAD", 6, d);
    run<2>("CE: VIMNMX x2", 2, d);
    run<3>("CE: VIMNMX + 2 IMAD", 3, d);
    run<4>("2 funnel shifts + xor", 3, d);
    run<5>("IMAD x2", 2, d);
    run<6>("IMAD.HI x2 (+add)", 3, d);
    run<7>("LOP3 x2", 2, d);
    run<8>("LOP3 + IMAD", 2, d);
    run<9>("funnel via IMAD.HI + IMAD", 2, d);
    run<11>("6 LOP3 (+add)", 7, d);
    run<10>("6 LOP3 + 2 IMAD.HI (+2 add)", 10, d);
    run<12>("6 LOP3 + 2 IMAD.WIDE (+adds)", 10, d);
    run<13>("6 LOP3 + 4 IMAD (+2 add)", 12, d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
