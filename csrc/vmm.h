// cuMem VMM symmetric allocations + NVLS multicast objects (see vmm.cc).
#pragma once
#include <pybind11/pybind11.h>

namespace tfos {
void bind_vmm(pybind11::module_& m);
}
