// TEST INFRASTRUCTURE — restatement of gym3's libenv C interface (gym3==0.3.3 `gym3/libenv.h`,
// pinned by /root/reference/environment.yml:12; located by builder.py:83). gym3 is not in
// /root/reference nor installed here, so the header is reconstructed from its uses in
// vecgame.cpp:42-99,212-282,333-361,437-457 and vecoptions.cpp:4-54. The product-side copy of the
// same ABI is include/procgen_b200.h; tests assert the two agree on struct sizes/offsets.
#pragma once
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LIBENV_API __declspec(dllexport)
#else
#define LIBENV_API __attribute__((visibility("default")))
#endif

#define LIBENV_VERSION 1
#define LIBENV_MAX_NAME_LEN 128
#define LIBENV_MAX_NDIM 16

enum libenv_dtype {
    LIBENV_DTYPE_UNUSED = 0,
    LIBENV_DTYPE_UINT8 = 1,
    LIBENV_DTYPE_INT32 = 2,
    LIBENV_DTYPE_FLOAT32 = 3,
};

enum libenv_scalar_type {
    LIBENV_SCALAR_TYPE_UNUSED = 0,
    LIBENV_SCALAR_TYPE_REAL = 1,
    LIBENV_SCALAR_TYPE_DISCRETE = 2,
};

enum libenv_space_name {
    LIBENV_SPACE_UNUSED = 0,
    LIBENV_SPACE_OBSERVATION = 1,
    LIBENV_SPACE_ACTION = 2,
    LIBENV_SPACE_INFO = 3,
};

union libenv_value {
    uint8_t uint8;
    int32_t int32;
    float float32;
};

struct libenv_tensortype {
    char name[LIBENV_MAX_NAME_LEN];
    enum libenv_scalar_type scalar_type;
    enum libenv_dtype dtype;
    int shape[LIBENV_MAX_NDIM];
    int ndim;
    union libenv_value low;
    union libenv_value high;
};

struct libenv_option {
    char name[LIBENV_MAX_NAME_LEN];
    enum libenv_dtype dtype;
    int count;
    void *data;
};

struct libenv_options {
    struct libenv_option *items;
    int count;
};

struct libenv_buffers {
    void **ob;
    float *rew;
    uint8_t *first;
    void **info;
    void **ac;
};

typedef void libenv_env;

LIBENV_API int libenv_version();
LIBENV_API libenv_env *libenv_make(int num_envs, const struct libenv_options options);
LIBENV_API int libenv_get_tensortypes(libenv_env *handle, enum libenv_space_name name, struct libenv_tensortype *types);
LIBENV_API void libenv_set_buffers(libenv_env *handle, struct libenv_buffers *bufs);
LIBENV_API void libenv_observe(libenv_env *handle);
LIBENV_API void libenv_act(libenv_env *handle);
LIBENV_API void libenv_close(libenv_env *handle);

#ifdef __cplusplus
}
#endif
