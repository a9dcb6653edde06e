#pragma once
#include "tensorflow/core/framework/op_kernel.h"  // syntax-check stub: everything lives in one header
