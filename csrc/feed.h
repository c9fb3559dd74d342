// Shared-memory pinned ring: the InputMode.SPARK data path.
//
// The reference moves every RDD row through a multiprocessing manager queue as
// an individually pickled proxy call (tensorflowonspark/TFSparkNode.py:500-502
// feeder side, tensorflowonspark/TFNode.py:278-300 consumer side).  Here the
// feeder task writes whole row blocks straight into a POSIX shared-memory ring
// that the training process has page-locked (cudaHostRegister), so a block
// goes feeder -> pinned host -> cudaMemcpyAsync on a side stream -> device with
// no per-row IPC and no intermediate copy.
#pragma once
#include <pybind11/pybind11.h>

namespace tfos {
void bind_feed(pybind11::module_& m);
}
