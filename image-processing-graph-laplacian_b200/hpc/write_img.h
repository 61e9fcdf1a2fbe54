/* Compatibility header: code written against the reference includes "write_img.h"; the declarations live in hpc_api.h. */
#include "hpc_api.h"
