#include "glc_internal.cuh"
using namespace glc;
extern "C" glc_status glc_flac_encode(glc_ctx *, const float *, uint64_t, uint32_t, uint16_t, uint8_t, uint8_t **, uint64_t *)
{ return set_error(GLC_ERR_UNSUPPORTED, "flac: not built yet"); }
extern "C" glc_status glc_flac_encode_batch(glc_ctx *, uint32_t, const float *const *, const uint64_t *, const uint32_t *, const uint16_t *, uint8_t, uint8_t **, uint64_t *)
{ return set_error(GLC_ERR_UNSUPPORTED, "flac: not built yet"); }
