// coeb_host.hpp -- host-side helpers shared by the C-ABI translation units (internal).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <string>

#include "../../include/coeb_types.h"

namespace coeb {

extern thread_local std::string g_last_error;

// Records a message for coeb_last_error() and returns `code`.
int fail(int code, const char* fmt, ...);

// COEB_OK if `device` exists and is sm_100 or newer; there is no CPU fallback.
int check_device(int device);

size_t select_smem_bytes(int max_nodes);

}  // namespace coeb

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::coeb::fail(COEB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
