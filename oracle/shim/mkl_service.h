#pragma once
/* oracle shim: see mkl.h */
#include "mkl.h"
