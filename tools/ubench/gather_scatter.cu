// Micro-benchmark: cost of random 4-byte gathers / scatters on B200 under different load/store
// flavours and locality windows.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gs gather_scatter.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ u32 hash32(u32 x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
// index for element l: random within a window of `win` elements that advances with l
__device__ __forceinline__ u32 idx_of(u32 l, u32 n, u32 win){ u32 base = (u32)(((u64)l / win) * win); u32 w = (base + win <= n) ? win : (n - base); return base + hash32(l) % w; }

template<int MODE> __global__ void k_gather(const u32* __restrict__ tab, u32* __restrict__ out, u32 n, u32 win){
  u32 stride = gridDim.x*blockDim.x;
  for (u32 l = blockIdx.x*blockDim.x+threadIdx.x; l < n; l += stride){
    u32 i = idx_of(l, n, win); u32 v;
    if (MODE==0) v = __ldg(tab+i);
    else if (MODE==1) asm volatile("ld.global.u32 %0,[%1];":"=r"(v):"l"(tab+i));
    else if (MODE==2) asm volatile("ld.global.cg.u32 %0,[%1];":"=r"(v):"l"(tab+i));
    else if (MODE==3) asm volatile("ld.global.L1::no_allocate.u32 %0,[%1];":"=r"(v):"l"(tab+i));
    else if (MODE==4) asm volatile("ld.global.cv.u32 %0,[%1];":"=r"(v):"l"(tab+i));
    else if (MODE==5) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u32 %0,[%1];":"=r"(v):"l"(tab+i));
    else asm volatile("ld.global.lu.u32 %0,[%1];":"=r"(v):"l"(tab+i));
    out[l] = v;
  }
}
template<int MODE> __global__ void k_scatter(u32* __restrict__ tab, u32 n, u32 win){
  u32 stride = gridDim.x*blockDim.x;
  for (u32 l = blockIdx.x*blockDim.x+threadIdx.x; l < n; l += stride){
    u32 i = idx_of(l, n, win);
    if (MODE==0) tab[i] = l;
    else if (MODE==1) asm volatile("st.global.cg.u32 [%0],%1;"::"l"(tab+i),"r"(l));
    else if (MODE==2) asm volatile("st.global.L1::no_allocate.u32 [%0],%1;"::"l"(tab+i),"r"(l));
    else if (MODE==3) asm volatile("st.global.wt.u32 [%0],%1;"::"l"(tab+i),"r"(l));
    else { u64* t8 = (u64*)tab; t8[i>>1] = ((u64)l<<32)|l; }   // 8-byte scatter
  }
}
template<typename F> float timeit(F f, int reps=3){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); cudaEventRecord(a); for(int i=0;i<reps;i++) f(); cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,a,b); return ms/reps; }

int main(int argc, char** argv){
  u32 n = 1u<<28; if (argc>1) n = (u32)atol(argv[1]);
  size_t lim=0; cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit: %zu\n", lim);
  if (argc>2){ cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[2])); cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity); printf("set -> %s, now %zu\n", cudaGetErrorString(e), lim); }
  u32 *tab,*out; CK(cudaMalloc(&tab,(size_t)n*4)); CK(cudaMalloc(&out,(size_t)n*4)); CK(cudaMemset(tab,1,(size_t)n*4));
  int blocks = 148*8, thr = 512;
  u32 wins[] = {n, 1u<<25, 1u<<23, 1u<<21, 1u<<16};
  const char* gn[] = {"__ldg","ld.global","ld.cg","ld.L1::no_allocate","ld.cv","ld.nc.L2::64B","ld.lu"};
  for (u32 win : wins){
    printf("window %u elements (%.1f MB)\n", win, win*4.0/1e6);
    float ms;
    ms = timeit([&]{k_gather<0><<<blocks,thr>>>(tab,out,n,win);}); printf("  gather %-22s %8.3f ms  %6.1f Gelem/s\n", gn[0], ms, n/ms/1e6);
    ms = timeit([&]{k_gather<1><<<blocks,thr>>>(tab,out,n,win);}); printf("  gather %-22s %8.3f ms  %6.1f Gelem/s\n", gn[1], ms, n/ms/1e6);
    ms = timeit([&]{k_gather<2><<<blocks,thr>>>(tab,out,n,win);}); printf("  gather %-22s %8.3f ms  %6.1f Gelem/s\n", gn[2], ms, n/ms/1e6);
    ms = timeit([&]{k_gather<3><<<blocks,thr>>>(tab,out,n,win);}); printf("  gather %-22s %8.3f ms  %6.1f Gelem/s\n", gn[3], ms, n/ms/1e6);
    ms = timeit([&]{k_gather<4><<<blocks,thr>>>(tab,out,n,win);}); printf("  gather %-22s %8.3f ms  %6.1f Gelem/s\n", gn[4], ms, n/ms/1e6);
    ms = timeit([&]{k_gather<5><<<blocks,thr>>>(tab,out,n,win);}); printf("  gather %-22s %8.3f ms  %6.1f Gelem/s\n", gn[5], ms, n/ms/1e6);
    ms = timeit([&]{k_scatter<0><<<blocks,thr>>>(tab,n,win);}); printf("  scatter %-21s %8.3f ms  %6.1f Gelem/s\n", "st", ms, n/ms/1e6);
    ms = timeit([&]{k_scatter<1><<<blocks,thr>>>(tab,n,win);}); printf("  scatter %-21s %8.3f ms  %6.1f Gelem/s\n", "st.cg", ms, n/ms/1e6);
    ms = timeit([&]{k_scatter<3><<<blocks,thr>>>(tab,n,win);}); printf("  scatter %-21s %8.3f ms  %6.1f Gelem/s\n", "st.wt", ms, n/ms/1e6);
    ms = timeit([&]{k_scatter<4><<<blocks,thr>>>(tab,n,win);}); printf("  scatter %-21s %8.3f ms  %6.1f Gelem/s\n", "st 8B", ms, n/ms/1e6);
  }
  return 0;
}
