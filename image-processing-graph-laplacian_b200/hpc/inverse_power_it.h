/* Compatibility header: code written against the reference includes "inverse_power_it.h"; the declarations live in hpc_api.h. */
#include "hpc_api.h"
