"""Generates the committed fixtures under tests/golden/ (run in the authoring container, where
/root/reference exists; the GPU box only reads the .npz files).

For each case the reference ships (examples/opf/14bus/case.raw|.rop, examples/opf/case300.m,
examples/opt/acopf/case3.m):
  <case>_data.npz         the PowersenseData tables produced by powersense_b200.create_PowersenseData
                          (restating src/opf/PowersenseData.jl:107-214 on the parsed file)
  <case>_<FORM>.npz       a seeded evaluation point (x, mu) with g / J / H and both sparsity structures from
                          oracle/jump_oracle.py -- the expression-tree restatement of JuMP's published
                          algorithm -- which pins oracle/ps_oracle.c and the CUDA kernels.
The reference's own evaluator (JuMP 0.21.10) cannot run here (no Julia): parity is UNPINNED against it."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import powersense_b200 as ps  # noqa: E402
from powersense_b200 import synth  # noqa: E402
from oracle import jump_oracle as jo  # noqa: E402

REF = "/root/reference/examples/"
CASES = {
    "case14": [REF + "opf/14bus/case.raw", REF + "opf/14bus/case.rop"],
    "case300": REF + "opf/case300.m",
    "case3": REF + "opt/acopf/case3.m",
}
ARRAY_FIELDS = ("Vmax Vmin Pmax Pmin Qmax Qmin bmax bmin Imax Gf Bf Gt Bt gf bf gt bt br Pd Qd Gl Bl gen_ptr gen_idx "
                "sij_ptr sij_idx sji_ptr sji_idx sh_ptr sh_idx lps c0 c1 c2").split()


def save_data(path, d):
    out = {k: getattr(d, k) for k in ARRAY_FIELDS}
    out["pgh_flat"] = np.concatenate([np.asarray(a, dtype=np.float64) for a in d.pgh]) if d.pgh else np.zeros(0)
    out["cgh_flat"] = np.concatenate([np.asarray(a, dtype=np.float64) for a in d.cgh]) if d.cgh else np.zeros(0)
    out["pgh_len"] = np.array([len(a) for a in d.pgh], dtype=np.int64)
    out["meta"] = np.array([d.nbus, d.ngen, d.nbr, d.nbss, d.cost_order, d.slack], dtype=np.int64)
    out["source_type"] = np.array(d.source_type)
    np.savez_compressed(path, **out)


def main():
    for name, path in CASES.items():
        d = ps.create_PowersenseData(path)
        save_data(os.path.join(HERE, f"{name}_data.npz"), d)
        for F in ps.ALL_FORMULATIONS:
            m = jo.build_model(d, F.name)
            x, _, _, lay = synth.scenario_batch(d, F, 1, seed=1234 + F.code)
            mu = np.random.default_rng(99 + F.code).normal(size=len(m.rows))
            g, J, H = m.evaluate(x[0], mu)
            jr, jc, _ = m.jacobian_structure()
            hr, hc, _ = m.hessian_structure()
            lo, hi = m.bounds()
            np.savez_compressed(os.path.join(HERE, f"{name}_{F.name}.npz"), x=x[0], mu=mu, g=g, J=J, H=H, jrow=jr, jcol=jc,
                                hrow=hr, hcol=hc, lo=lo, hi=hi, n_var=np.int64(m.n_var))
            print(name, F.name, len(g), len(J), len(H))


if __name__ == "__main__":
    main()
