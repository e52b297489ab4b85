// Reference include path kept for source compatibility; everything lives in op/hotpath.h.
#pragma once
#include "op/hotpath.h"
