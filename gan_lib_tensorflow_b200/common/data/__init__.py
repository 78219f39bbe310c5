"""Input side of the reference scripts (`common/data`): only the CIFAR-10 batch reader is mirrored -- it is what
`lib.data.cifar10.load` hands to the training loops of SNGAN/gan_cifar_resnet.py and ACGAN/train.py."""
