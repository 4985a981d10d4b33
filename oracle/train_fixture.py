"""TEST INFRASTRUCTURE ONLY. Train the SfNeural architecture (NNManager.create_net, nn_manager.py:277-298) on seeded
synthetic boards so that the tests have a set of REALISTIC weights next to the seeded Glorot init: the reference's trained
model does not ship (cvconf.py:58 is a download URL), and random weights exercise neither peaked softmax outputs nor a
meaningful stone stream. Writes tests/golden/sfneural_trained.npz (flat float32 blob, camkifu_b200.weights layout).

    python -m oracle.train_fixture [--boards 700] [--epochs 6]        (authoring container, CPU, a few minutes)

Inputs are raw 0..255 like the reference's (nn_cache.py:49-50: no scaling); labels follow NNManager.compute_label.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def make_data(n_boards, seed):
    import cv2
    from camkifu_b200 import synth
    from camkifu_b200.plugins import subregion
    from oracle import oracle as O
    xs, ys = [], []
    done = 0
    k = 0
    while done < n_boards:
        n = min(8, n_boards - done)
        frames, mtx, truth, _ = synth.make_clip(seed + k, n, 240 + 60 * (k % 3), 320 + 80 * (k % 3))
        k += 1
        for f, st in zip(frames, truth):
            g = cv2.warpPerspective(f, mtx, (380, 380))
            xs.append(O.c_nn_gather(g))
            lab = np.empty(100, np.int64)
            for i in range(10):
                for j in range(10):
                    rs, re, cs, ce = subregion(i, j)
                    sq = st[rs:re, cs:ce].astype(np.int64).ravel()       # row-major 2x2, least significant first
                    lab[i * 10 + j] = int(sq[0] + 3 * sq[1] + 9 * sq[2] + 27 * sq[3])
            ys.append(lab)
        done += n
    return np.concatenate(xs), np.concatenate(ys)


def main():
    import torch
    import torch.nn as nn
    from camkifu_b200 import weights
    ap = argparse.ArgumentParser()
    ap.add_argument("--boards", type=int, default=700)
    ap.add_argument("--epochs", type=int, default=6)
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    X, Y = make_data(args.boards, seed=5000)
    Xv, Yv = make_data(40, seed=9000)
    print("data: %d patches (%.0f s)" % (len(X), time.time() - t0), flush=True)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.c1, self.c2 = nn.Conv2d(3, 32, 5), nn.Conv2d(32, 32, 5)
            self.c3, self.c4 = nn.Conv2d(32, 90, 3), nn.Conv2d(90, 90, 3)
            self.f5, self.f6 = nn.Linear(3240, 160), nn.Linear(160, 81)

        def forward(self, x, train=False):
            h = torch.relu(self.c2(torch.relu(self.c1(x))))
            h = torch.max_pool2d(h, 2)
            h = torch.relu(self.c4(torch.relu(self.c3(h))))
            h = torch.max_pool2d(h, 2)
            h = h.permute(0, 2, 3, 1).reshape(x.shape[0], 3240)       # Keras Flatten: (H, W, C)
            h = torch.relu(self.f5(h))
            if train:
                h = torch.dropout(h, 0.3, True)
            return self.f6(h)

    net = Net()
    with torch.no_grad():                                             # raw 0..255 inputs: start the first layer small
        net.c1.weight.mul_(1.0 / 64)
    opt = torch.optim.Adam(net.parameters(), lr=4e-4)
    Xt = torch.from_numpy(X)
    Yt = torch.from_numpy(Y)

    def batches(Xa, Ya, bs, shuffle):
        idx = torch.randperm(len(Xa)) if shuffle else torch.arange(len(Xa))
        for s in range(0, len(Xa), bs):
            b = idx[s:s + bs]
            yield Xa[b].to(torch.float32).permute(0, 3, 1, 2), Ya[b]

    for ep in range(args.epochs):
        net.train()
        tot, n = 0.0, 0
        for xb, yb in batches(Xt, Yt, 256, True):
            opt.zero_grad()
            loss = torch.nn.functional.cross_entropy(net(xb, train=True), yb)
            loss.backward()
            opt.step()
            tot += float(loss) * len(yb)
            n += len(yb)
        net.eval()
        with torch.no_grad():
            acc = np.mean([float((net(xb).argmax(1) == yb).float().mean()) for xb, yb in
                           batches(torch.from_numpy(Xv), torch.from_numpy(Yv), 500, False)])
        print("epoch %d: loss %.4f  val patch accuracy %.4f  (%.0f s)" % (ep, tot / n, acc, time.time() - t0), flush=True)
        if ep == args.epochs - 2:
            for g_ in opt.param_groups:
                g_["lr"] = 1e-4

    sd = net.state_dict()
    parts = [sd["c1.weight"].permute(2, 3, 1, 0), sd["c1.bias"], sd["c2.weight"].permute(2, 3, 1, 0), sd["c2.bias"],
             sd["c3.weight"].permute(2, 3, 1, 0), sd["c3.bias"], sd["c4.weight"].permute(2, 3, 1, 0), sd["c4.bias"],
             sd["f5.weight"].t(), sd["f5.bias"], sd["f6.weight"].t(), sd["f6.bias"]]
    flat = weights.from_keras_weights([p.contiguous().numpy() for p in parts])
    out = os.path.join(ROOT, "tests", "golden", "sfneural_trained.npz")
    np.savez_compressed(out, params=flat, val_patch_accuracy=np.float64(acc), boards=np.int64(args.boards))
    print("wrote", out, flat.shape, "val accuracy", acc)


if __name__ == "__main__":
    main()
