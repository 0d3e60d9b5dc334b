// Sparse LU (symbolic on host, numeric on device) -- internal interface.
#pragma once
#include "wae_internal.h"
