"""I/O edge of the SVM path: the MNIST IDX reader and the one-vs-all driver of examples/mnistsvm.m.

``readmnist`` follows readMNIST / trimDigits / normalizePixValue (mnistsvm.m:188-259) byte for byte:
big-endian int32 headers (2051 images / 2049 labels), count / offset checks with the reference's error
texts, uint8 pixels read row by row, a 4-pixel border trimmed (28x28 -> 20x20), pixels divided by 255.
``digit_rows`` is the flattening of mnistsvm.m:96-107 (``reshape(im', 1, 400)``: row-major).
``mnistsvm`` mirrors the driver (mnistsvm.m:60-168): sample with replacement, train ten one-vs-all
classifiers with the hinge and the '01' loss, report the reference's error percentages -- the ten hinge
classifiers run as ONE class batch on the device (linearsvm_onevsall)."""
from __future__ import annotations

import os
import struct

import numpy as np

from .errorcheck import MatlabError
from .solvers.linearsvm import linearsvm, linearsvm_onevsall


def readmnist(img_file, label_file, read_digits, offset=0, border=4):
    """[imgs, labels] = readMNIST(imgFile, labelFile, readDigits, offset), mnistsvm.m:188-241.
    imgs: (h - 2*border) x (w - 2*border) x readDigits float64 in [0, 1]; labels: readDigits float64."""
    with open(img_file, "rb") as f:
        header, = struct.unpack(">i", f.read(4))                            # fopen(..., 'r', 'b'): big endian
        if header != 2051:
            raise MatlabError("Invalid image file header")
        count, = struct.unpack(">i", f.read(4))
        if count < read_digits + offset:
            raise MatlabError("Trying to read too many digits")
        h, w = struct.unpack(">ii", f.read(8))
        if offset > 0:
            f.seek(w * h * offset, os.SEEK_CUR)
        raw = np.frombuffer(f.read(w * h * read_digits), dtype=np.uint8)
    if raw.size != w * h * read_digits:
        raise MatlabError("Trying to read too many digits")
    imgs = raw.reshape(read_digits, h, w).transpose(1, 2, 0).astype(np.float64)     # imgs(y,:,i) = row y of digit i
    with open(label_file, "rb") as f:
        header, = struct.unpack(">i", f.read(4))
        if header != 2049:
            raise MatlabError("Invalid label file header")
        count, = struct.unpack(">i", f.read(4))
        if count < read_digits + offset:
            raise MatlabError("Trying to read too many digits")
        if offset > 0:
            f.seek(offset, os.SEEK_CUR)
        labels = np.frombuffer(f.read(read_digits), dtype=np.uint8).astype(np.float64)
    imgs = imgs[border:h - border, border:w - border, :]                    # trimDigits, :243-250
    return imgs / 255.0, labels                                             # normalizePixValue, :252-259


def digit_rows(imgs):
    """mnistsvm.m:100-107: row i = reshape(imgs(:,:,i)', 1, h*w) -- the digit flattened row by row."""
    h, w, n = imgs.shape
    return np.asfortranarray(imgs.transpose(2, 0, 1).reshape(n, h * w))


def one_vs_all_labels(labels, nclass=10):
    """trainForDigit's relabelling (mnistsvm.m:170-178) for every digit at once: column k is +1 where
    the label equals k, -1 elsewhere."""
    labels = np.asarray(labels).reshape(-1)
    return np.asfortranarray(np.where(labels[:, None] == np.arange(nclass)[None, :], 1.0, -1.0))


def mnistsvm(C=0.5, rho=1.0, testsubsets=1000, trainsubsets=6000, data_dir="MNIST", engine=None, losses=("hinge", "01"),
             counts=(10000, 60000), quiet=False):
    """examples/mnistsvm.m:35-168.  Returns (result, X): result is the 10 x 4 table of error percentages
    [hinge train, 0-1 train, hinge test, 0-1 test] (:144-155; a column stays NaN for a loss not in
    ``losses``), X maps a loss to its 400 x 10 matrix of classifiers."""
    test_im, test_lab = readmnist(os.path.join(data_dir, "t10k-images.idx3-ubyte"),
                                  os.path.join(data_dir, "t10k-labels.idx1-ubyte"), counts[0], 0)
    train_im, train_lab = readmnist(os.path.join(data_dir, "train-images.idx3-ubyte"),
                                    os.path.join(data_dir, "train-labels.idx1-ubyte"), counts[1], 0)
    test_all, train_all = digit_rows(test_im), digit_rows(train_im)
    # datasample(X, k): k rows drawn uniformly WITH replacement (:112-113)
    ti = np.random.randint(0, test_all.shape[0], size=int(testsubsets))
    ri = np.random.randint(0, train_all.shape[0], size=int(trainsubsets))
    test_vec, test_lab = test_all[ti], test_lab[ti]
    train_vec, train_lab = np.asfortranarray(train_all[ri]), train_lab[ri]
    m, n = test_lab.size, train_lab.size
    ELL, TELL = one_vs_all_labels(train_lab), one_vs_all_labels(test_lab)
    options = {"rho": float(rho), "maxiters": 500, "algorithm": "fast"}     # trainForDigit, :180-182
    result = np.full((10, 4), np.nan)
    X = {}
    if "hinge" in losses:       # ten linearsvm(D, ell_k, C, options) calls of :183 as one class batch
        res = linearsvm_onevsall(train_vec, ELL, C, options, engine=engine)
        X["hinge"] = np.stack([r["xopt"] for r in res], axis=1)
    if "01" in losses:          # :184-185, one call per digit (the 0-1 prox usually trips the H-norm return)
        cols = []
        for k in range(10):
            r = linearsvm(train_vec, ELL[:, k], C, dict(options, lossfunction="01", history=0), engine=engine)
            cols.append(r["xopt"] if "xopt" in r else np.full(train_vec.shape[1], np.nan))
        X["01"] = np.stack(cols, axis=1)
    for j, loss in enumerate(("hinge", "01")):
        if loss in X:                                                       # :144-155
            result[:, j] = np.sum((1 - ELL * (train_vec @ X[loss])) > 0, axis=0) / n * 100.0
            result[:, 2 + j] = np.sum((1 - TELL * (test_vec @ X[loss])) > 0, axis=0) / m * 100.0
    if not quiet:
        print("\nError Percentages:\n")
        print("Digit\tHinge (Train)\t0-1 (Train)\tHinge (Test)\t0-1 (Test)")
        for k in range(10):
            print("%d\t\t%2.4f\t\t\t%2.4f\t\t%2.4f\t\t\t%2.4f" % (k, result[k, 0], result[k, 1], result[k, 2], result[k, 3]))
    return result, X
