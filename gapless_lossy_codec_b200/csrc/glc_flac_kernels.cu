// placeholder, replaced below
#include "glc_internal.cuh"
