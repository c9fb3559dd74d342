"""Shared pieces of the MNIST examples: one training loop that runs the torch module on CPU
executors (gloo) and the native sm_100a trainer on B200 executors (fused all-reduce + SGD)."""
import time


def make_trainer(ctx, batch_size, lr):
  """Returns (step_fn(images_uint8[B,784], labels[B]) -> loss, export_fn(dir, is_chief), desc)."""
  import numpy as np
  import torch
  from tensorflowonspark_b200 import compat
  from tensorflowonspark_b200.models import mnist, simple
  sig = {"serving_default": {"inputs": {"image": "image"},
                             "outputs": {"logits": "logits", "prediction": "prediction"},
                             "input_shapes": {"image": [-1, 784]}}}
  use_gpu = bool(ctx.gpus) and torch.cuda.is_available()
  if use_gpu:
    torch.cuda.set_device(0)
    comm = ctx.symmetric_comm() if ctx.world_size > 1 else None
    net = mnist.MnistTrainer(batch=batch_size, device="cuda:0", lr=lr, comm=comm)
    if comm is not None:
      comm.broadcast("weights", root=0)
      comm.broadcast("aux32", root=0)
    xh = torch.zeros(batch_size, 784, dtype=torch.float32).pin_memory()
    yh = torch.zeros(batch_size, dtype=torch.int32).pin_memory()

    def step(images, labels):
      xh.copy_(torch.from_numpy(np.asarray(images, dtype=np.float32) / 255.0))
      yh.copy_(torch.from_numpy(np.asarray(labels, dtype=np.int32)))
      return float(net.train_step(xh, yh).item())

    def export(export_dir, is_chief):
      compat.export_saved_model(net, export_dir, is_chief, signatures=sig)

    return step, export, "native sm_100a trainer, world {}".format(ctx.world_size)

  if ctx.world_size > 1:
    ctx.init_process_group(backend="gloo")
  torch.manual_seed(1234)
  model = mnist.MnistCNN()
  opt = torch.optim.SGD(model.parameters(), lr=lr)

  def step(images, labels):
    x = torch.from_numpy(np.asarray(images, dtype=np.float32) / 255.0)
    y = torch.from_numpy(np.asarray(labels, dtype=np.int64))
    loss = torch.nn.functional.cross_entropy(model(x), y)
    opt.zero_grad()
    loss.backward()
    simple.allreduce_mean_grads(model, ctx.world_size)
    opt.step()
    return float(loss.detach())

  def export(export_dir, is_chief):
    compat.export_saved_model(model, export_dir, is_chief, signatures=sig)

  return step, export, "torch CPU module + gloo, world {}".format(ctx.world_size)


class StepTimer(object):

  def __init__(self, every=100):
    self.every, self.t0, self.n = every, time.time(), 0

  def tick(self, step, loss, batch):
    self.n += 1
    if self.n % self.every == 0:
      dt = time.time() - self.t0
      print("step {:6d} loss {:.4f}  {:.0f} images/s".format(step, loss, self.every * batch / dt))
      self.t0 = time.time()
