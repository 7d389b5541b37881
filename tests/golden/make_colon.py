"""Generates tests/golden/colon.npz from the reference's own data set and sampler (run in the build container,
where /root/reference exists; the GPU box only sees the committed fixture).

  * parses /root/reference/data/colon.rda (bzip2'd RDX2 / XDR serialisation) with a minimal reader:
    colon$X 62 x 2000 gene expressions, colon$y 0/1 labels (R/data.R, data/colon.rda);
  * forms htlr_fit()'s inputs the way R/core.r:103-146 does (std(): column median / sample sd, src/utils.cpp:36-48;
    intercept column; ybase / ymat), zero initial coefficients and the sigmasbt of R/core.r:192-193;
  * runs the reference's htlr_fit_helper (oracle/_ref/libhtlr_ref.so = its sources compiled verbatim) with
    injected normal / uniform / gamma streams and records the chain.
"""
import bz2
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


class RdaReader:
    def __init__(self, raw):
        assert raw[:5] == b"RDX2\n" and raw[5:7] == b"X\n", "expected an XDR RDX2 file"
        self.b, self.at = raw, 7
        self.refs = []
        self.version = [self.int() for _ in range(3)]

    def int(self):
        v = struct.unpack_from(">i", self.b, self.at)[0]
        self.at += 4
        return v

    def item(self):
        flags = self.int()
        typ, has_attr, has_tag = flags & 0xFF, bool(flags & 0x200), bool(flags & 0x400)
        if typ == 254:      # NILVALUE
            return None
        if typ == 255:      # REFSXP
            return self.refs[(flags >> 8) - 1]
        if typ == 1:        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 9:        # CHARSXP
            n = self.int()
            if n < 0:
                return None
            s = self.b[self.at:self.at + n].decode("latin1")
            self.at += n
            return s
        if typ == 2:        # LISTSXP (pairlist): attr, tag, car, cdr
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.int()
                typ, has_attr, has_tag = flags & 0xFF, bool(flags & 0x200), bool(flags & 0x400)
                if typ == 254:
                    return out
                assert typ == 2, typ
        if typ in (13, 10):  # INTSXP / LGLSXP
            n = self.int()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.at).astype(np.int64)
            self.at += 4 * n
        elif typ == 14:     # REALSXP
            n = self.int()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.at).astype(np.float64)
            self.at += 8 * n
        elif typ == 16:     # STRSXP
            n = self.int()
            v = [self.item() for _ in range(n)]
        elif typ == 19:     # VECSXP
            n = self.int()
            v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError(f"SEXP type {typ}")
        attrs = dict(self.item()) if has_attr else {}
        if "dim" in attrs and isinstance(v, np.ndarray):
            v = v.reshape(tuple(int(d) for d in attrs["dim"]), order="F")
        if typ == 19 and "names" in attrs:
            return dict(zip(attrs["names"], v))
        return v


def load_colon():
    raw = bz2.decompress(open("/root/reference/data/colon.rda", "rb").read())
    top = dict(RdaReader(raw).item())
    colon = top["colon"]
    return np.asarray(colon["X"], dtype=np.float64), np.asarray(colon["y"]).astype(np.int64)


def main():
    from htlr_b200 import synth
    from oracle import oracle as O
    X, y = load_colon()
    assert X.shape == (62, 2000) and set(np.unique(y)) == {0, 1} and int(y.sum()) == 40
    Xs = synth.std_median_sd(X)
    Xa, ymat, ybase, K = synth.make_inputs(Xs, y)
    p = X.shape[1]
    deltas, sig = synth.init_state(p, K, init="zero", seed=77)
    args = dict(p=p, K=K, n=62, X=Xa, ymat=ymat, ybase=ybase, ptype="t", alpha=1.0, s=-10.0, eta=0.0, iters_rmc=12,
                iters_h=12, thin=1, leap_L=50, leap_L_h=5, leap_step=0.3, hmc_sgmcut=0.05, deltas=deltas,
                sigmasbt=sig, keep_warmup_hist=True, silence=1, legacy=False)
    g = np.random.default_rng(2024)
    T = 24
    st = dict(normals=g.standard_normal(T * (p + 1) * K), uniforms=g.random(T), gammas=g.gamma((1.0 + K) / 2, size=T * p))
    ref = O.ref_fit(**args, **st)
    out = os.path.join(ROOT, "tests", "golden", "colon.npz")
    np.savez_compressed(out, X=X.astype(np.float32) if np.array_equal(X.astype(np.float32).astype(np.float64), X) else X,
                        y=y, sigmasbt=sig, stream_seed=2024,
                        mcdeltas_nz=ref["mcdeltas"][np.abs(ref["mcdeltas"]).max(axis=(1, 2)) > 0],
                        nz_rows=np.nonzero(np.abs(ref["mcdeltas"]).max(axis=(1, 2)) > 0)[0],
                        mcloglike=ref["mcloglike"], mcuvar=ref["mcuvar"], mchmcrej=ref["mchmcrej"],
                        mcsigmasbt_last=ref["mcsigmasbt"][:, -1], mcvardeltas_last=ref["mcvardeltas"][:, -1],
                        consumed=ref["consumed"])
    print("wrote", out, os.path.getsize(out), "bytes; uvar", ref["mcuvar"][-5:], "rej", ref["mchmcrej"].mean(),
          "consumed", ref["consumed"])


if __name__ == "__main__":
    main()
