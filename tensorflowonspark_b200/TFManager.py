"""Per-executor IPC hub between Spark feeder tasks and the training process.

Keeps the reference contract (tensorflowonspark/TFManager.py:40-83): ``start(authkey, queues,
mode)`` launches a manager process exposing named ``JoinableQueue``s through
``get_queue(name)`` plus a small key/value store through ``get``/``set`` (the ``'state'`` key
walks ``running`` -> ``terminating`` -> ``stopped``); ``connect(address, authkey)`` attaches from
another process.  ``mode='local'`` listens on a unix socket (workers), ``'remote'`` on TCP so the
driver can reach ps / evaluator nodes.

Bulk row data does not have to go through these queues any more: when a shared-memory ring is
registered (``set('ring', {...})``) feeders write row blocks into it and only post a small
``marker.RingBlock`` descriptor here, which preserves the queue's join()/task_done() accounting.
"""
import logging
import multiprocessing
from multiprocessing.managers import BaseManager

logger = logging.getLogger(__name__)


class TFManager(BaseManager):
  """Manager type; the served callables are registered in :func:`start` / :func:`connect`."""


# state owned by the manager *server* process (populated before it forks off in start())
_hub = {"queues": {}, "kv": {}}


def _hub_queue(name):
  try:
    return _hub["queues"][name]
  except KeyError:
    return None


def _hub_get(key):
  return _hub["kv"].get(key)


def _hub_set(key, value):
  _hub["kv"][key] = value


def start(authkey, queues, mode="local"):
  """Create the queues, then start the manager server.

  Args:
    authkey: bytes used to authenticate connections.
    queues: names of the JoinableQueues to create (e.g. ``['input', 'output', 'error']``).
    mode: ``'local'`` (unix socket) or ``'remote'`` (TCP on an ephemeral port).
  Returns:
    the started manager; ``mgr.address`` is what peers pass to :func:`connect`.
  """
  _hub["queues"] = {name: multiprocessing.JoinableQueue() for name in queues}
  _hub["kv"] = {}
  TFManager.register("get_queue", callable=_hub_queue)
  TFManager.register("get", callable=_hub_get)
  TFManager.register("set", callable=_hub_set)
  if mode == "remote":
    mgr = TFManager(address=("", 0), authkey=authkey)
  else:
    mgr = TFManager(authkey=authkey)
  mgr.start()
  # The server process is shut down when its owning manager object is finalised; later tasks
  # on the same (reused) python worker rebind TFSparkNode.mgr to a *client* connection, so the
  # owner must be pinned here for the life of the executor process.
  _owned[:] = [mgr]
  logger.debug("TFManager started in %s mode at %s", mode, mgr.address)
  return mgr


_owned = []


def connect(address, authkey):
  """Attach to a running manager at ``address`` (unix path or (host, port))."""
  TFManager.register("get_queue")
  TFManager.register("get")
  TFManager.register("set")
  if isinstance(address, list):
    address = tuple(address)
  m = TFManager(address, authkey=authkey)
  m.connect()
  return m
