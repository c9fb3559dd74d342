"""ResNet-56 / CIFAR-10 on Spark executors (InputMode.TENSORFLOW): argv is passed through to
``resnet_cifar_main.main_fun`` which parses its own flags on every node (reference:
examples/resnet/resnet_cifar_spark.py:17-22).

  python examples/resnet/resnet_cifar_spark.py --cluster_size 8 --use_synthetic_data --train_steps 200
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import resnet_cifar_main  # noqa: E402

if __name__ == "__main__":
  from tensorflowonspark_b200 import TFCluster
  from tensorflowonspark_b200._spark import SparkConf, SparkContext

  args, rem = resnet_cifar_main.define_flags().parse_known_args(sys.argv[1:])
  conf = SparkConf().setAppName("resnet_cifar") \
      .set("spark.executor.instances", str(args.cluster_size)) \
      .set("spark.executor.resource.gpu.amount", "1").set("spark.task.resource.gpu.amount", "1")
  sc = SparkContext(conf=conf)
  cluster = TFCluster.run(sc, resnet_cifar_main.main_fun, sys.argv, args.cluster_size, num_ps=0,
                          input_mode=TFCluster.InputMode.TENSORFLOW, master_node="chief")
  cluster.shutdown()
  sc.stop()
