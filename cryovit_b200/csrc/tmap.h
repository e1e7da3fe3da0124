// Host-side helpers shared by the C-ABI translation units: error slot, TMA tensor-map encoder
// (driver entry point fetched at run time so the library links against cudart only and can be
// dlopen'ed on a box without libcuda for the symbol-export test).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvit {

// Error codes returned across the C ABI (0 == ok). Mirrored in include/cryovit_b200.h.
enum : int {
  CVIT_OK = 0,
  CVIT_ERR_INVALID = -1,   // bad shape / pointer / alignment
  CVIT_ERR_CUDA = -2,      // a CUDA runtime call or launch failed
  CVIT_ERR_DRIVER = -3,    // driver entry point (tensor-map encode) unavailable or failed
  CVIT_ERR_UNSUPPORTED = -4
};

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int num_sms();

enum class TmapDtype { BF16, F16, F32 };

// Encode a tiled tensor map. dims/strides innermost-first; strides in BYTES for dims 1..rank-1.
// swizzle_bytes in {0,32,64,128}. Returns CVIT_OK or an error code.
int encode_tmap(CUtensorMap* out, TmapDtype dt, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

}  // namespace cvit
