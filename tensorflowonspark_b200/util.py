"""Host utilities (reference: tensorflowonspark/util.py:21-94)."""
import errno
import logging
import os
import socket
import subprocess

logger = logging.getLogger(__name__)

EXECUTOR_ID_FILE = "executor_id"


def single_node_env(num_gpus=1, worker_index=-1, nodes=None):
  """Prepare the environment of an independent single-node process: expand the Hadoop
  classpath once and pin ``CUDA_VISIBLE_DEVICES`` to this worker's share of the host's GPUs.

  Args:
    num_gpus: GPUs wanted by this process (0 = CPU only).
    worker_index: global worker index, or -1 for "any free GPUs".
    nodes: optional list of host addresses of all workers; when given, the GPU slot is the
      worker's index *among the workers on the same host* (reference util.py:31-49).
  """
  from . import gpu_info
  if "HADOOP_PREFIX" in os.environ and "TFOS_CLASSPATH_UPDATED" not in os.environ:
    try:
      extra = subprocess.check_output(
          [os.path.join(os.environ["HADOOP_PREFIX"], "bin", "hadoop"), "classpath", "--glob"])
      os.environ["CLASSPATH"] = os.environ.get("CLASSPATH", "") + os.pathsep + extra.decode().strip()
    except Exception as e:  # hadoop is optional on a single box
      logger.debug("hadoop classpath expansion skipped: %s", e)
    os.environ["TFOS_CLASSPATH_UPDATED"] = "1"

  if num_gpus > 0 and gpu_info.is_gpu_available():
    slot = worker_index
    if nodes and worker_index >= 0:
      me = nodes[worker_index]
      slot = [i for i, addr in enumerate(nodes) if addr == me].index(worker_index)
    gpus = gpu_info.get_gpus(num_gpus, slot)
    logger.info("Using gpu(s): %s", gpus)
    os.environ["CUDA_VISIBLE_DEVICES"] = gpus
  else:
    logger.info("Using CPU")
    os.environ["CUDA_VISIBLE_DEVICES"] = ""


def get_ip_address():
  """IP address of this host as seen on the default route; loopback-safe."""
  s = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
  try:
    s.connect(("8.8.8.8", 1))  # UDP connect sends nothing; it only selects the outgoing interface
    return s.getsockname()[0]
  except socket.error:
    try:
      return socket.gethostbyname(socket.getfqdn())
    except socket.error:
      return "127.0.0.1"
  finally:
    s.close()


def usable_cpus():
  """CPUs this process may really use: affinity mask capped by the cgroup CPU quota."""
  try:
    n = len(os.sched_getaffinity(0))
  except (AttributeError, OSError):
    n = os.cpu_count() or 1
  for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try:
      with open(path) as f:
        parts = f.read().split()
      if path.endswith("cpu.max"):
        if parts[0] != "max":
          n = min(n, max(1, int(parts[0]) // int(parts[1])))
      else:
        quota = int(parts[0])
        if quota > 0:
          with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
            n = min(n, max(1, quota // int(f.read())))
      break
    except (IOError, OSError, ValueError, IndexError):
      continue
  return max(1, n)


def limit_intra_op_threads(local_nodes):
  """Default OMP/MKL/torch thread counts for a node process that shares the host with
  ``local_nodes - 1`` others (one process per GPU): an unbounded intra-op pool per process
  oversubscribes the box - or its container quota - by an order of magnitude.  Explicit
  ``OMP_NUM_THREADS`` wins."""
  if os.environ.get("OMP_NUM_THREADS"):
    return int(os.environ["OMP_NUM_THREADS"])
  n = max(1, min(8, usable_cpus() // max(1, 2 * local_nodes)))
  os.environ["OMP_NUM_THREADS"] = str(n)
  os.environ.setdefault("MKL_NUM_THREADS", str(n))
  try:
    import sys
    if "torch" in sys.modules:
      sys.modules["torch"].set_num_threads(n)
  except Exception:
    pass
  return n


def find_in_path(path, file):
  """First ``<dir>/<file>`` that exists for the directories of a PATH-like string, else False."""
  for p in path.split(os.pathsep):
    candidate = os.path.join(p, file)
    if os.path.exists(candidate) and os.path.isfile(candidate):
      return candidate
  return False


def write_executor_id(num):
  """Record this executor's id in its working directory - the join key between the long-running
  node task and the feeder tasks that land on the same executor later."""
  with open(EXECUTOR_ID_FILE, "w") as f:
    f.write(str(num))


def read_executor_id():
  """Read back the id written by :func:`write_executor_id` (raises with the usual causes if absent)."""
  try:
    with open(EXECUTOR_ID_FILE, "r") as f:
      return int(f.read())
  except (IOError, OSError) as e:
    if getattr(e, "errno", None) not in (None, errno.ENOENT):
      raise
    raise Exception(
        "No executor_id file found on this node ({}). Check that: 1) the cluster has exactly one "
        "task slot per executor (spark.task.cpus == spark.executor.cores), 2) dynamic allocation "
        "is off, 3) the number of executors equals the cluster size, 4) python workers are "
        "reused (spark.python.worker.reuse).".format(os.getcwd()))
