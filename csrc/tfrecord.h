// Native TFRecord framing (masked CRC32C) and tf.train.Example codec.
// Replaces the JVM pieces of the reference: lib/tensorflow-hadoop-1.0-SNAPSHOT.jar
// (TFRecordFileInputFormat/OutputFormat, Crc32C) and the Example conversion in
// src/main/scala/com/yahoo/tensorflowonspark/DFUtil.scala:119-258.
#pragma once
#include <pybind11/pybind11.h>

namespace tfos {
void bind_tfrecord(pybind11::module_& m);
}
