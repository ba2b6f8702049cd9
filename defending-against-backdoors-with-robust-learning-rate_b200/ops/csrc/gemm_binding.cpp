#include <torch/extension.h>
#include "gemm.h"
void register_gemm_bindings(py::module_& m) { (void)m; }
