// Include overlay for the reference's utils/cuda/errors.cuh (:9-30): the same three macros.  Host code that is built
// against librsfusion.so has no CUDA calls of its own to check, so they only evaluate their argument.
#pragma once
#ifndef CUDA_SAFE_CALL
#define CUDA_SAFE_CALL(err) ((void)(err))
#define CUDA_CHECK_ERROR ((void)0)
#define CUDA_STREAM_CHECK_ERROR(stream) ((void)(stream))
#endif
