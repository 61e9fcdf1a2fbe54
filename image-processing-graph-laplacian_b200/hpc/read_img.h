/* Compatibility header: code written against the reference includes "read_img.h"; the declarations live in hpc_api.h. */
#include "hpc_api.h"
