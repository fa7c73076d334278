// Micro-benchmark 2: what does ONE random 4-byte access cost on B200, in time and in DRAM bytes, under load / store
// flavours that take different paths through L2 (plain, strong, atomic, texture, wide, full-sector)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o g2 gather2.cu
// Run under: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv ./g2
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ u32 hash32(u32 x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template<int MODE> __global__ void k_gather(u32* __restrict__ tab, u32* __restrict__ out, u32 n, u32 mask, cudaTextureObject_t tex){
  u32 stride = gridDim.x*blockDim.x;
  for (u32 l = blockIdx.x*blockDim.x+threadIdx.x; l < n; l += stride){
    u32 i = hash32(l) & mask; u32 v;
    if (MODE==0) v = __ldg(tab+i);
    else if (MODE==1) v = atomicAdd(tab+i, 0u);
    else if (MODE==2) asm volatile("ld.relaxed.gpu.global.u32 %0,[%1];":"=r"(v):"l"(tab+i));
    else if (MODE==3) { u64 pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;":"=l"(pol)); asm volatile("ld.global.nc.L2::cache_hint.u32 %0,[%1],%2;":"=r"(v):"l"(tab+i),"l"(pol)); }
    else if (MODE==4) v = tex1Dfetch<u32>(tex, (int)i);
    else if (MODE==5) { unsigned short b; asm volatile("ld.global.nc.u8 %0,[%1];":"=h"(b):"l"((const unsigned char*)(tab+i))); v = b; }
    else if (MODE==6) { uint4 q = __ldg((const uint4*)(tab + (i & ~3u))); v = q.x ^ q.y ^ q.z ^ q.w; }
    else { uint4 q0 = __ldg((const uint4*)(tab + (i & ~7u))); uint4 q1 = __ldg((const uint4*)(tab + (i & ~7u) + 4)); v = q0.x ^ q0.w ^ q1.x ^ q1.w; }
    out[l] = v;
  }
}
template<int MODE> __global__ void k_scatter(u32* __restrict__ tab, u32 n, u32 mask){
  u32 stride = gridDim.x*blockDim.x;
  for (u32 l = blockIdx.x*blockDim.x+threadIdx.x; l < n; l += stride){
    u32 i = hash32(l) & mask;
    if (MODE==0) tab[i] = l;
    else if (MODE==1) atomicAdd(tab+i, l);                    // RED
    else if (MODE==2) atomicExch(tab+i, l);                   // ATOM (result unused)
    else if (MODE==3) asm volatile("st.relaxed.gpu.global.u32 [%0],%1;"::"l"(tab+i),"r"(l));
    else if (MODE==4) { u64 pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;":"=l"(pol)); asm volatile("st.global.L2::cache_hint.u32 [%0],%1,%2;"::"l"(tab+i),"r"(l),"l"(pol)); }
    else if (MODE==5) ((unsigned char*)tab)[(size_t)i*4] = (unsigned char)l;
    else if (MODE==6) *(uint4*)(tab + (i & ~3u)) = make_uint4(l,l,l,l);          // 16 B, aligned
    else { uint4 q = make_uint4(l,l,l,l); *(uint4*)(tab + (i & ~7u)) = q; *(uint4*)(tab + (i & ~7u) + 4) = q; }  // one whole 32 B sector
  }
}
template<typename F> float timeit(F f, int reps=2){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); CK(cudaDeviceSynchronize()); cudaEventRecord(a); for(int i=0;i<reps;i++) f(); cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms,a,b); return ms/reps; }

int main(int argc, char** argv){
  u32 n = 1u<<27; u32 tabn = 1u<<28;  // 2^27 accesses into a 1 GiB table
  if (argc>1) tabn = (u32)atol(argv[1]);
  u32 *tab,*out; CK(cudaMalloc(&tab,(size_t)tabn*4)); CK(cudaMalloc(&out,(size_t)n*4)); CK(cudaMemset(tab,1,(size_t)tabn*4));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab; rd.res.linear.desc = cudaCreateChannelDesc<u32>(); rd.res.linear.sizeInBytes = (size_t)(tabn > (1u<<27) ? (1u<<27) : tabn)*4;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType; cudaTextureObject_t tex = 0; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  int blocks = 148*8, thr = 512; u32 mask = tabn - 1; float ms;
  const char* gn[] = {"ldg","atomicAdd(+0)","ld.relaxed.gpu","ld evict_first","tex1Dfetch(512MiB)","ld.u8","ld 16B","ld 2x16B sector"};
  const char* sn[] = {"st","red.add","atom.exch","st.relaxed.gpu","st evict_first","st.u8","st 16B","st 32B sector"};
#define G(M) ms = timeit([&]{k_gather<M><<<blocks,thr>>>(tab,out,n,(M==4)?(mask>>1):mask,tex);}); printf("gather  %-20s %8.3f ms  %6.1f G/s\n", gn[M], ms, n/ms/1e6);
#define S(M) ms = timeit([&]{k_scatter<M><<<blocks,thr>>>(tab,n,mask);}); printf("scatter %-20s %8.3f ms  %6.1f G/s\n", sn[M], ms, n/ms/1e6);
  G(0) G(1) G(2) G(3) G(4) G(5) G(6) G(7)
  S(0) S(1) S(2) S(3) S(4) S(5) S(6) S(7)
  return 0;
}
