"""Run N independent single-node instances of a function on the executors.

Parity with tensorflowonspark/TFParallel.py:17-74: an (optional) barrier stage guarantees that
all instances are scheduled together and lets each one see the addresses of the others, so it
can take its own slot of the host's GPUs; the function receives a bare ``TFNodeContext`` with
``worker_num`` / ``executor_id`` / ``num_workers`` / ``defaultFS`` and shards its own input
(reference examples/mnist/keras/mnist_inference.py:41-45).
"""
import logging

from . import TFSparkNode, util
from ._spark import BarrierTaskContext

logger = logging.getLogger(__name__)


def run(sc, map_fn, tf_args, num_executors, use_barrier=True):
  """Run ``map_fn(tf_args, ctx)`` once on each of ``num_executors`` executors.

  Args:
    sc: SparkContext.
    map_fn: user function.
    tf_args: arguments for ``map_fn`` (``num_gpus`` is honoured if present, default 1).
    num_executors: number of independent instances.
    use_barrier: schedule them as one barrier stage (all-or-nothing).
  Returns:
    the concatenated results of the instances (whatever iterables ``map_fn`` returns).
  """
  default_fs = sc._jsc.hadoopConfiguration().get("fs.defaultFS")
  if default_fs.startswith("file://") and len(default_fs) > 7 and default_fs.endswith("/"):
    default_fs = default_fs[:-1]

  def _run(it):
    worker_num = None
    for i in it:
      worker_num = i
    nodes = None
    if use_barrier:
      infos = BarrierTaskContext.get().getTaskInfos()
      nodes = [t.address.split(":")[0] for t in infos]
    num_gpus = 1
    if "num_gpus" in tf_args:
      num_gpus = int(tf_args.num_gpus if hasattr(tf_args, "num_gpus") else tf_args["num_gpus"])
    util.single_node_env(num_gpus=num_gpus, worker_index=worker_num, nodes=nodes)
    ctx = TFSparkNode.TFNodeContext()
    ctx.defaultFS = default_fs
    ctx.worker_num = worker_num
    ctx.executor_id = worker_num
    ctx.num_workers = len(nodes) if nodes is not None else num_executors
    ctx.rank, ctx.world_size = worker_num, ctx.num_workers
    out = map_fn(tf_args, ctx)
    return out if out is not None else []

  node_rdd = sc.parallelize(list(range(num_executors)), num_executors)
  if use_barrier:
    return node_rdd.barrier().mapPartitions(_run).collect()
  return node_rdd.mapPartitions(_run).collect()
