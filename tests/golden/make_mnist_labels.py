"""Writes tests/golden/mnist_train_labels_6000.npy: the first 6000 labels of the reference's own label file
examples/MNIST/train-labels.idx1-ubyte (IDX1, big-endian header 2049 / 60000; examples/mnistsvm.m:188-241).
The GPU box has no /root/reference, so the labels the one-vs-all SVM tests use travel as this small fixture.
    python tests/golden/make_mnist_labels.py"""
import os
import struct

import numpy as np

SRC = "/root/reference/examples/MNIST/train-labels.idx1-ubyte"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    raw = open(SRC, "rb").read()
    magic, count = struct.unpack(">ii", raw[:8])
    assert magic == 2049 and count == 60000
    lab = np.frombuffer(raw[8:8 + 6000], dtype=np.uint8).copy()
    np.save(os.path.join(HERE, "mnist_train_labels_6000.npy"), lab)
    print("wrote", lab.size, "labels; class counts", np.bincount(lab))
