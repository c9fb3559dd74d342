"""Restart-from-checkpoint recovery for a whole training job.

The reference detects failures and aborts (error queue -> feeder task /
``shutdown()`` -> ``sc.stop()`` + ``sys.exit(1)``, tensorflowonspark/TFCluster.py:179-183); getting
the job going again is left to whoever submitted it, with TensorFlow's implicit resume from
``model_dir`` (examples/mnist/estimator/mnist_spark.py:94-97).  ``run_with_restarts`` is that
outer loop as a library call: it runs the driver body, and when the cluster reports a failed
node (an exception or the ``SystemExit`` of ``TFCluster.shutdown``) it tears the Spark context
down, builds a fresh one - new executors, new node processes, new process group - and runs the
body again.  The body's ``map_fun`` is expected to resume from the newest checkpoint
(``utils.checkpoint.latest_checkpoint`` - atomic writes guarantee it is complete), so an attempt
only repeats the steps since the last save.  Not elastic: the restarted job has the same shape.

  def job(sc, attempt):
    cluster = TFCluster.run(sc, main_fun, args, n, 0, input_mode=TFCluster.InputMode.TENSORFLOW)
    cluster.shutdown()

  attempts = recovery.run_with_restarts(lambda: SparkContext(conf=conf), job, max_restarts=2)
"""
import logging
import time

logger = logging.getLogger(__name__)


class JobFailed(Exception):
  """Every attempt failed; ``causes`` holds the exception (or SystemExit) of each attempt."""

  def __init__(self, causes):
    super(JobFailed, self).__init__("job failed {} time(s); last: {!r}".format(len(causes), causes[-1]))
    self.causes = causes


def run_with_restarts(make_context, job, max_restarts=2, backoff_s=1.0, on_failure=None):
  """Run ``job(sc, attempt)`` until it returns; restart it on a fresh context when it fails.

  Args:
    make_context: ``() -> SparkContext`` - called once per attempt (a failed cluster has usually
      stopped its context, and its executors may hold dead node processes).
    job: the driver body; anything it raises - including the ``SystemExit`` with which
      ``TFCluster.shutdown`` / ``reservation.await_reservations`` leave a failed application -
      counts as a failed attempt.  ``KeyboardInterrupt`` is never swallowed.
    max_restarts: restarts after the first attempt (0 = plain run).
    backoff_s: pause before attempt k is ``backoff_s * k`` seconds.
    on_failure: ``(attempt, exc) -> None`` hook between attempts (alerting, clearing state).

  Returns the number of attempts used (1 = no failure).  Raises :class:`JobFailed` when the last
  allowed attempt failed too.
  """
  causes = []
  for attempt in range(max_restarts + 1):
    sc = make_context()
    try:
      job(sc, attempt)
      try:
        sc.stop()
      except Exception:
        pass
      if causes:
        logger.info("job completed on attempt %d after %d failure(s)", attempt + 1, len(causes))
      return attempt + 1
    except KeyboardInterrupt:
      raise
    except BaseException as e:   # noqa: B902 - SystemExit is how a failed cluster leaves the driver
      causes.append(e)
      logger.error("attempt %d failed: %r", attempt + 1, e)
      try:
        sc.cancelAllJobs()
      except Exception:
        pass
      try:
        sc.stop()
      except Exception:
        pass
      if on_failure is not None:
        on_failure(attempt, e)
      if attempt < max_restarts:
        time.sleep(backoff_s * (attempt + 1))
  raise JobFailed(causes)
