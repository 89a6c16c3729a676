"""Regenerates tests/golden/ from the reference checkout (run in the build container only; /root/reference does not
exist on the GPU box).  Data files are the reference's bundled fixtures (`data/normal.*`, `data/multivariate.*`,
`data/covariates.txt`); the known-answer numbers are transcribed from the reference's own documentation
(`docs/src/man/examples.md:230-268`, MendelIHT v1.4.1, `iht("normal", 7, Normal, covariates="covariates.txt")`).
"""
import json
import os
import shutil

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for f in ("normal.bed", "multivariate.bed", "covariates.txt", "normal_true_beta.txt"):
        shutil.copyfile(os.path.join(REF, "data", f), os.path.join(HERE, f))
    fam = np.loadtxt(os.path.join(REF, "data", "normal.fam"))
    np.savetxt(os.path.join(HERE, "normal_y.txt"), fam[:, 5], fmt="%.17g")
    mfam = np.loadtxt(os.path.join(REF, "data", "multivariate.fam"))
    np.savetxt(os.path.join(HERE, "multivariate_y.txt"), mfam[:, 5:7], fmt="%.17g")
    docs = {
        "source": "docs/src/man/examples.md:230-268",
        "call": 'iht("normal", 7, Normal, covariates="covariates.txt", phenotypes=6)',
        "n": 1000, "p": 10000, "k": 7,
        "logl": [-1403.6085154464329, -1397.922430744325, -1397.8812223841496, -1397.8807476657355,
                 -1397.8807416751808],
        "tol": [0.8141937613701785, 0.017959863148623176, 0.001989846075839033, 0.00016446741159857614,
                2.0482155566893502e-5],
        "backtracks": [0, 0, 0, 0, 0],
        "iter": 5,
        "final_logl": -1397.8807416751808,
        "pve": 0.8343751445053728,
        "support_1based": [3137, 4246, 4717, 6290, 7755, 8375, 9415],
        "beta_6sig": [0.424376, 0.52343, 0.922857, -0.677832, -0.542983, -0.792813, -2.17998],
        "c_6sig": [1.65223, 0.749865],
    }
    with open(os.path.join(HERE, "docs_trace_normal_k7.json"), "w") as f:
        json.dump(docs, f, indent=1)


if __name__ == "__main__":
    main()
