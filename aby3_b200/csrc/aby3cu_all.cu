// aby3cu_all.cu -- the library is compiled as ONE translation unit so that the
// __constant__ AES table and the inline device functions need no relocatable
// device code.  Build: see aby3_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
#include "abi.cu"
#include "elementwise.cu"
#include "gemm_imad.cu"
#include "gemm_tc.cu"
#include "binary.cu"
#include "sharedot.cu"
#include "batched.cu"
#include "sgd_fused.cu"
