#pragma once
#include <math.h>
#define CUDART_INF ((double)INFINITY)
