"""Node bring-up and the GPU-allocation matrix (scenarios of reference tests/test_TFSparkNode.py:
38-190), driven by calling the node closure directly against a real reservation server."""
import os
from unittest import mock

import pytest

from tensorflowonspark_b200 import TFManager, TFSparkNode, gpu_info, reservation


def _run_node(fn, tf_args, template={"worker": [0]}, executor=0, **patches):
  server = reservation.Server(1)
  addr = server.start()
  meta = {"id": 12345, "cluster_template": template, "num_executors": 1, "default_fs": "file://",
          "working_dir": os.getcwd(), "server_addr": addr, "release_port": True}
  TFSparkNode.TFSparkNode.mgr = None
  cwd = os.getcwd()
  try:
    node = TFSparkNode.run(fn, tf_args, meta, False, None, ["input", "output", "error"], False)
    node([executor])
  finally:
    os.chdir(cwd)
    server.stop()
    if os.path.exists("executor_id"):
      os.remove("executor_id")
    m = TFSparkNode.TFSparkNode.mgr
    if m is not None:
      try:
        m.shutdown()
      except Exception:
        pass
    TFSparkNode.TFSparkNode.mgr = None


def test_role_and_context():
  seen = {}

  def fn(args, ctx):
    seen.update(job=ctx.job_name, idx=ctx.task_index, ex=ctx.executor_id, spec=ctx.cluster_spec,
                nw=ctx.num_workers, rank=ctx.rank, world=ctx.world_size,
                tf_config=os.environ.get("TF_CONFIG"), master=os.environ.get("MASTER_ADDR"),
                sock=ctx.tmp_socket)

  with mock.patch.object(gpu_info, "is_gpu_available", return_value=False):
    _run_node(fn, {}, template={"chief": [0]})
  assert seen["job"] == "chief" and seen["idx"] == 0 and seen["ex"] == 0
  assert list(seen["spec"]) == ["chief"] and seen["nw"] == 1
  assert seen["rank"] == 0 and seen["world"] == 1 and seen["master"]
  assert '"type": "chief"' in seen["tf_config"] and seen["sock"] is None


def test_gpu_unavailable_but_requested():
  with mock.patch.object(gpu_info, "is_gpu_available", return_value=False), \
       mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=False):
    with pytest.raises(Exception, match="none is available"):
      _run_node(lambda a, c: None, {"num_gpus": 1})


def test_gpu_available_default_one():
  got = {}

  def fn(args, ctx):
    got["visible"] = os.environ["CUDA_VISIBLE_DEVICES"]
    got["gpus"] = ctx.gpus

  with mock.patch.object(gpu_info, "is_gpu_available", return_value=True), \
       mock.patch.object(gpu_info, "get_gpus", return_value=["0"]) as gg, \
       mock.patch.object(gpu_info, "_inventory", side_effect=Exception("no nvml")), \
       mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=False):
    _run_node(fn, {})
  assert got == {"visible": "0", "gpus": ["0"]}
  assert gg.call_args_list[0] == mock.call(1, 0, format=gpu_info.AS_LIST)


def test_peers_stay_visible_in_first_mode():
  got = {}

  def fn(args, ctx):
    got["visible"] = os.environ["CUDA_VISIBLE_DEVICES"]

  inv = ([(i, "GPU-%d" % i) for i in range(4)], set())
  with mock.patch.object(gpu_info, "is_gpu_available", return_value=True), \
       mock.patch.object(gpu_info, "get_gpus", return_value=["2"]), \
       mock.patch.object(gpu_info, "_inventory", return_value=inv), \
       mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=False):
    _run_node(fn, {"num_gpus": 1})
  assert got["visible"] == "2,0,1,3"  # assigned GPU is device 0, peers remain mappable


def test_spark_resource_api_wins():
  got = {}

  def fn(args, ctx):
    got["gpus"] = ctx.gpus

  res = {"gpu": mock.Mock(addresses=["3", "1"])}
  tctx = mock.Mock()
  tctx.resources.return_value = res
  with mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=True), \
       mock.patch.object(TFSparkNode.TaskContext, "get", return_value=tctx), \
       mock.patch.object(gpu_info, "_inventory", side_effect=Exception("x")), \
       mock.patch.object(gpu_info, "get_gpus") as gg:
    _run_node(fn, {"num_gpus": 1})
  assert got["gpus"] == ["3"] and not gg.called


def test_kubernetes_never_guesses():
  got = {}

  def fn(args, ctx):
    got["visible"] = os.environ["CUDA_VISIBLE_DEVICES"]

  with mock.patch.dict(os.environ, {"SPARK_EXECUTOR_POD_IP": "1.2.3.4"}), \
       mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=False), \
       mock.patch.object(gpu_info, "get_gpus") as gg:
    _run_node(fn, {})
    assert got["visible"] == "" and not gg.called
    with pytest.raises(Exception, match="Kubernetes"):
      _run_node(fn, {"num_gpus": 1})


def test_second_node_task_on_live_executor_is_rejected():
  with mock.patch.object(gpu_info, "is_gpu_available", return_value=False):
    server = reservation.Server(1)
    addr = server.start()
    meta = {"id": 777, "cluster_template": {"worker": [0]}, "num_executors": 1,
            "default_fs": "file://", "working_dir": os.getcwd(), "server_addr": addr}
    TFSparkNode.TFSparkNode.mgr = None
    node = TFSparkNode.run(lambda a, c: None, {}, meta, False, None, ["input", "error"], False)
    node([0])
    with pytest.raises(Exception, match="already started"):
      node([0])
    TFSparkNode.TFSparkNode.mgr.set("state", "stopped")
    server.stop()
    TFManager._owned[0].shutdown()
    TFSparkNode.TFSparkNode.mgr = None
    os.remove("executor_id")


def test_gpu_index_is_per_host_position_in_the_cluster_spec():
  """Several nodes on one host must take disjoint GPUs: the slot handed to gpu_info is the node's
  position among the nodes OF ITS HOST in the (mocked) cluster spec, not the executor id
  (reference tests/test_TFSparkNode.py:89-106: worker:1 of 1.1.1.1 -> index 2)."""
  spec = {"chief": ["1.1.1.1:2222"], "worker": ["1.1.1.1:2223", "1.1.1.1:2224", "2.2.2.2:2222"]}
  calls = []

  def fake_get_gpus(n, index, format=None):  # noqa: A002
    calls.append((n, index, format))
    return ["0"]

  with mock.patch.object(gpu_info, "is_gpu_available", return_value=True), \
       mock.patch.object(gpu_info, "get_gpus", side_effect=fake_get_gpus), \
       mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=False), \
       mock.patch.dict(os.environ, {}, clear=False):
    os.environ.pop("SPARK_EXECUTOR_POD_IP", None)
    for job, idx, want in (("chief", 0, 0), ("worker", 0, 1), ("worker", 1, 2), ("worker", 2, 0)):
      TFSparkNode._get_gpus({"num_gpus": 1}, 7, cluster_spec=spec, job_name=job, task_index=idx)
      assert calls[-1] == (1, want, gpu_info.AS_LIST), (job, idx, calls[-1])


def test_spark_resource_api_without_gpu_resources_falls_back_to_gpu_info():
  """reference test_gpu_spark_fallback (:133-152): Spark 3 is there but the job was submitted
  without spark.executor.resource.gpu.* - the node still finds free GPUs through nvidia-smi."""
  got = {}

  def fn(args, ctx):
    got["gpus"], got["visible"] = ctx.gpus, os.environ["CUDA_VISIBLE_DEVICES"]

  tctx = mock.Mock()
  tctx.resources.return_value = {}
  with mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=True), \
       mock.patch.object(TFSparkNode.TaskContext, "get", return_value=tctx), \
       mock.patch.object(gpu_info, "is_gpu_available", return_value=True), \
       mock.patch.object(gpu_info, "get_gpus", return_value=["5"]) as gg:
    _run_node(fn, {"num_gpus": 1})
  assert got["gpus"] == ["5"] and gg.called and got["visible"].split(",")[0] == "5"


def test_no_gpu_anywhere_defaults_to_cpu_but_an_explicit_request_fails():
  """reference test_gpu_spark_unavailable_default / _but_requested (:154-190)."""
  got = {}

  def fn(args, ctx):
    got["gpus"], got["visible"] = ctx.gpus, os.environ["CUDA_VISIBLE_DEVICES"]

  tctx = mock.Mock()
  tctx.resources.return_value = {}
  with mock.patch.object(TFSparkNode, "_has_spark_resource_api", return_value=True), \
       mock.patch.object(TFSparkNode.TaskContext, "get", return_value=tctx), \
       mock.patch.object(gpu_info, "is_gpu_available", return_value=False), \
       mock.patch.object(gpu_info, "get_gpus") as gg:
    _run_node(fn, {})
    assert got == {"gpus": [], "visible": ""} and not gg.called
    with pytest.raises(Exception, match="requested but none"):
      _run_node(fn, {"num_gpus": 1})
