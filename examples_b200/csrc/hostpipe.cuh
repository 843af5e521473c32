// Host <-> device transfers of the stateless drop-in entry points (hostpipe.cu; internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace atlj {
bool host_ptr_is_pinned(const void *p);  // cudaHostAlloc / cudaHostRegister memory
int hostpipe_threads();                  // host threads of the staging copies
// host -> device in stream order on st; a pageable source is consumed when the call returns
int hostpipe_h2d(cudaStream_t st, void *dst_dev, const void *src_host, size_t bytes);
// device -> host; the data is in dst_host (stream synchronised) when the call returns
int hostpipe_d2h(cudaStream_t st, void *dst_host, const void *src_dev, size_t bytes);
}  // namespace atlj
