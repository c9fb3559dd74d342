"""Model zoo of the reference workloads: MNIST CNN, ResNet (CIFAR 56 / ImageNet 50),
segmentation U-Net; plus small models used by the pipeline tests."""
