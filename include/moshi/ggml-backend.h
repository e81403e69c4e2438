/* ggml-backend.h — shim.  The reference's public header includes <ggml-backend.h> only for the `ggml_backend *` arguments of
 * moshi_alloc (include/moshi/moshi.h:28).  This build has no ggml: the type is opaque, the arguments are ignored (the CUDA
 * device is chosen with moshi_alloc_b200 or the MOSHI_B200_DEVICE environment variable). */
#pragma once
struct ggml_backend;
typedef struct ggml_backend *ggml_backend_t;
