/* Compatibility header: code written against the reference includes "sampling.h"; the declarations live in hpc_api.h. */
#include "hpc_api.h"
