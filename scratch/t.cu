__global__ void k(float* out, const int* idx, const float* v){
  __shared__ float acc[256];
  acc[threadIdx.x]=0; __syncthreads();
  atomicAdd(&acc[idx[threadIdx.x]], v[threadIdx.x]);
  asm volatile("red.shared.add.f32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&acc[idx[threadIdx.x+32]])), "f"(v[threadIdx.x]));
  __syncthreads(); out[threadIdx.x]=acc[threadIdx.x];
}
