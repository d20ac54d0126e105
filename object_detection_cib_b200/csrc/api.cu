// Version / error-string entry points of libkodb.
#include "kodb_common.cuh"

extern "C" int kodb_version(void) { return 200; }

extern "C" const char* kodb_error_string(int code) {
    switch (code) {
        case KODB_OK: return "ok";
        case KODB_ERR_INVALID_ARGUMENT: return "invalid argument";
        case KODB_ERR_CAPACITY: return "capacity exceeded (workspace or per-image target list too small)";
        case KODB_ERR_LAUNCH: return "CUDA launch error";
        case KODB_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}
