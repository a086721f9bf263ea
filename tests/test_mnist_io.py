"""The I/O edge of the SVM path: MNIST IDX reader parity with examples/mnistsvm.m:188-259 (big-endian
headers 2051 / 2049, offset, 4-pixel trim, /255, row-major flattening) on IDX files written by the test,
and -- on the GPU -- the one-vs-all driver fed by those files."""
import os
import struct

import numpy as np
import pytest

from admm_project_b200 import MatlabError
from admm_project_b200.mnist import digit_rows, mnistsvm, one_vs_all_labels, readmnist


def write_idx(tmp, name, images, labels, magic=(2051, 2049)):
    n, h, w = images.shape
    ip, lp = os.path.join(tmp, name + "-images.idx3-ubyte"), os.path.join(tmp, name + "-labels.idx1-ubyte")
    with open(ip, "wb") as f:
        f.write(struct.pack(">iiii", magic[0], n, h, w))
        f.write(images.astype(np.uint8).tobytes())
    with open(lp, "wb") as f:
        f.write(struct.pack(">ii", magic[1], n))
        f.write(labels.astype(np.uint8).tobytes())
    return ip, lp


def matlab_readmnist(images, labels, read_digits, offset):
    """readMNIST + trimDigits + normalizePixValue written out literally (loops and 1-based slices)."""
    n, h, w = images.shape
    imgs = np.zeros((h, w, read_digits))
    for i in range(read_digits):
        for y in range(h):
            imgs[y, :, i] = images[offset + i, y, :]
    border = 4
    out = np.zeros((h - 2 * border, w - 2 * border, read_digits))
    for i in range(read_digits):
        out[:, :, i] = imgs[border:h - border, border:w - border, i]
    for i in range(read_digits):
        out[:, :, i] = out[:, :, i] / 255.0
    return out, labels[offset:offset + read_digits].astype(np.float64)


def test_reader_matches_the_reference_reader(tmp_path):
    rs = np.random.RandomState(0)
    images = rs.randint(0, 256, size=(37, 28, 28))
    labels = rs.randint(0, 10, size=37)
    ip, lp = write_idx(str(tmp_path), "t", images, labels)
    for count, offset in [(37, 0), (10, 5), (1, 36)]:
        imgs, lab = readmnist(ip, lp, count, offset)
        ref_imgs, ref_lab = matlab_readmnist(images, labels, count, offset)
        assert imgs.shape == (20, 20, count) and np.array_equal(imgs, ref_imgs) and np.array_equal(lab, ref_lab)
        rows = digit_rows(imgs)
        for i in range(count):                                      # reshape(im', 1, 400)
            assert np.array_equal(rows[i], ref_imgs[:, :, i].T.reshape(-1, order="F"))
    assert imgs.max() <= 1.0 and imgs.dtype == np.float64


def test_reader_errors_are_the_references(tmp_path):
    images = np.zeros((5, 28, 28)); labels = np.arange(5)
    ip, lp = write_idx(str(tmp_path), "a", images, labels)
    with pytest.raises(MatlabError, match="Trying to read too many digits"):
        readmnist(ip, lp, 5, 1)
    bad_i, _ = write_idx(str(tmp_path), "b", images, labels, magic=(2049, 2049))
    with pytest.raises(MatlabError, match="Invalid image file header"):
        readmnist(bad_i, lp, 5, 0)
    _, bad_l = write_idx(str(tmp_path), "c", images, labels, magic=(2051, 2051))
    with pytest.raises(MatlabError, match="Invalid label file header"):
        readmnist(ip, bad_l, 5, 0)


def test_one_vs_all_labels_is_train_for_digits_relabelling():
    lab = np.array([3, 0, 9, 3, 1])
    E = one_vs_all_labels(lab)
    for d in range(10):                                             # mnistsvm.m:170-178
        ell = lab.astype(float).copy()
        for i in range(ell.size):
            ell[i] = 1 if lab[i] == d else -1
        assert np.array_equal(E[:, d], ell)


def synthetic_mnist(tmp, seed=0, ntrain=900, ntest=300):
    """28x28 uint8 digits whose class shows as a bright blob at a class-dependent position plus noise."""
    rs = np.random.RandomState(seed)

    def make(n):
        lab = rs.randint(0, 10, size=n)
        im = rs.randint(0, 60, size=(n, 28, 28))
        for i, d in enumerate(lab):
            r0, c0 = 5 + 3 * (d // 4), 5 + 4 * (d % 4)
            im[i, r0:r0 + 5, c0:c0 + 5] += 150
        return np.minimum(im, 255), lab
    tr, trl = make(ntrain)
    te, tel = make(ntest)
    os.makedirs(os.path.join(tmp, "MNIST"), exist_ok=True)
    write_idx(os.path.join(tmp, "MNIST"), "train", tr, trl)
    write_idx(os.path.join(tmp, "MNIST"), "t10k", te, tel)
    return os.path.join(tmp, "MNIST")


@pytest.mark.gpu
def test_mnistsvm_driver_on_idx_files(engine, tmp_path):
    import oracle
    d = synthetic_mnist(str(tmp_path))
    np.random.seed(4)
    result, X = mnistsvm(0.5, 1.0, 200, 1500, data_dir=d, engine=engine, losses=("hinge",), counts=(300, 900), quiet=True)
    assert X["hinge"].shape == (400, 10) and np.all(np.isnan(result[:, [1, 3]]))
    assert np.all(result[:, 0] < 25.0) and np.all(result[:, 2] < 35.0)      # margin violations, mnistsvm.m:144-151
    # the same draws through the oracle, digit by digit (trainForDigit, mnistsvm.m:170-186)
    np.random.seed(4)
    test_im, test_lab = readmnist(os.path.join(d, "t10k-images.idx3-ubyte"), os.path.join(d, "t10k-labels.idx1-ubyte"), 300)
    train_im, train_lab = readmnist(os.path.join(d, "train-images.idx3-ubyte"), os.path.join(d, "train-labels.idx1-ubyte"), 900)
    np.random.randint(0, 300, size=200)
    ri = np.random.randint(0, 900, size=1500)
    D, lab = digit_rows(train_im)[ri], train_lab[ri]
    for k in range(3):
        ref = oracle.linearsvm(D, np.where(lab == k, 1.0, -1.0), 0.5, {"rho": 1.0, "maxiters": 500, "history": 0})
        assert np.linalg.norm(X["hinge"][:, k] - ref["xopt"]) <= 1e-9 * np.linalg.norm(ref["xopt"])
