// Optional per-stage CUDA-event timing of a fused forward (bench.py's live roofline numbers); state lives in the context.
#pragma once
#include <cuda_runtime.h>
struct sps_ctx;
namespace sps {
void prof_begin(sps_ctx* c, cudaStream_t st);                     // first mark of a forward
void prof_mark(sps_ctx* c, const char* name, cudaStream_t st);    // closes the segment `name`
}
