// Probe: does a tiled TMA load of u8 with an innermost start coordinate that is not 16-byte aligned work on sm_100a?
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)
__device__ __forceinline__ uint32_t s32(const void* p){return (uint32_t)__cvta_generic_to_shared(p);}
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int z, uint8_t* out, int nbytes){
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar=(uint64_t*)(sm+4096);
  if(threadIdx.x==0){
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;":::"memory");
  }
  __syncthreads();
  if(threadIdx.x==0){
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"::"r"(s32(bar)),"r"(nbytes):"memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(s32(sm)),"l"(&tm),"r"(x),"r"(y),"r"(z),"r"(s32(bar)):"memory");
  }
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}"::"r"(s32(bar)):"memory");
  for(int i=threadIdx.x;i<nbytes;i+=blockDim.x) out[i]=sm[i];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(){
  const int W=176,H=144,P=2; std::vector<uint8_t> h(W*H*P); for(size_t i=0;i<h.size();i++) h[i]=(uint8_t)(i*7+ (i>>8));
  uint8_t* d; CK(cudaMalloc(&d,h.size())); CK(cudaMemcpy(d,h.data(),h.size(),cudaMemcpyHostToDevice));
  uint8_t* o; CK(cudaMalloc(&o,4096));
  void* p=nullptr; cudaDriverEntryPointQueryResult st; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",&p,cudaEnableDefault,&st)); Fn fn=(Fn)p;
  for(int bw : {48, 16}) for (int bh : {47, 16}) {
    CUtensorMap tm; cuuint64_t dims[3]={W,H,P}; cuuint64_t strides[2]={W,(cuuint64_t)W*H}; cuuint32_t box[3]={(cuuint32_t)bw,(cuuint32_t)bh,1}; cuuint32_t es[3]={1,1,1};
    CUresult r=fn(&tm,CU_TENSOR_MAP_DATA_TYPE_UINT8,3,d,dims,strides,box,es,CU_TENSOR_MAP_INTERLEAVE_NONE,CU_TENSOR_MAP_SWIZZLE_NONE,CU_TENSOR_MAP_L2_PROMOTION_L2_128B,CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode box %dx%d -> %d\n",bw,bh,(int)r); if(r) continue;
    for(int x : {16, 17, 18, 19, -16, -13, 150}) for (int y : {0, -15, 120}) {
      CK(cudaMemset(o,0xEE,4096));
      k<<<1,128,4096+64>>>(tm,x,y,1,o,bw*bh);
      cudaError_t e=cudaDeviceSynchronize();
      if(e!=cudaSuccess){printf("box %dx%d x=%d y=%d: %s\n",bw,bh,x,y,cudaGetErrorString(e)); return 2;}
      std::vector<uint8_t> g(bw*bh); CK(cudaMemcpy(g.data(),o,bw*bh,cudaMemcpyDeviceToHost));
      int bad=0; for(int r2=0;r2<bh;r2++) for(int c=0;c<bw;c++){int gx=x+c,gy=y+r2; uint8_t want=(gx>=0&&gx<W&&gy>=0&&gy<H)?h[(size_t)W*H+gy*W+gx]:0; if(g[r2*bw+c]!=want) bad++;}
      printf("box %dx%d x=%d y=%d bad=%d\n",bw,bh,x,y,bad);
    }
  }
  return 0;
}
