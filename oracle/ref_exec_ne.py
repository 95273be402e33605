"""Exec-oracle for the Norton-equivalent extraction: run the reference's
``Circuit Simulation/NE_from_sim.py`` (``NE`` below) on the simulation results that ship with the
reference (``smps.mat``: 4 harmonic supply frequencies x 2 magnitudes; ``circuit_sim.mat``: 5 x 3).

TEST INFRASTRUCTURE ONLY - used by ``oracle/make_golden_ne.py`` where ``/root/reference`` is
mounted.  Nothing is copied: the script text is read at run time, patched IN MEMORY and exec-ed:

* ``loadmat(device_name + "_" + str(f_max) + ".mat", ...)`` (NE:30) -> the given .mat path
  (the file the script names, ``SMPS_5050.mat``, is not in the repository);
* ``pd.Int64Index`` (NE:59; removed in pandas 2) -> ``pd.Index`` shim;
* ``Series.append`` / ``DataFrame.append`` (NE:117-118,127,135,166; removed in pandas 2) ->
  ``pd.concat`` shims installed for the duration of the run;
* ``np.zeros(...)`` used to pre-allocate DataFrames that later receive complex values
  (NE:62-66,145,157; pandas 3 refuses the silent float64 -> complex128 upcast older versions
  did) -> complex zeros;
* the two ``to_csv`` exports (NE:207-213) write into a scratch directory.
"""
from __future__ import annotations

import contextlib
import io
import os
import tempfile

import numpy as np
import pandas as pd

REF_ROOT = os.environ.get("HPF_REFERENCE_ROOT", "/root/reference")
NE_PATH = os.path.join(REF_ROOT, "Circuit Simulation", "NE_from_sim.py")


def available():
    return os.path.isfile(NE_PATH)


def run(mat_path):
    """-> globals dict of the exec-ed script (Y_N_c, I_N_c, Y_N_uc, I_N_uc, I_inj, freq, ...)."""
    src = open(NE_PATH).read()
    src = src.replace('loadmat(device_name + "_" + str(f_max) + ".mat",', 'loadmat(%r,' % mat_path)
    assert repr(mat_path) in src
    src = src.replace("np.zeros(", "_czeros(")
    g = {"__name__": "ne_from_sim_ref", "_czeros": lambda shape, dtype=None: np.zeros(shape, dtype=complex)}
    had_i64 = hasattr(pd, "Int64Index")
    if not had_i64:
        pd.Int64Index = lambda data, name=None, dtype=None: pd.Index(np.asarray(data), name=name, dtype=dtype)
    had_sa, had_da = hasattr(pd.Series, "append"), hasattr(pd.DataFrame, "append")
    if not had_sa:
        pd.Series.append = lambda self, other: pd.concat([self] + (list(other) if isinstance(other, (list, tuple)) else [other]))
    if not had_da:
        pd.DataFrame.append = lambda self, other: pd.concat([self] + (list(other) if isinstance(other, (list, tuple)) else [other]))
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="ne_ref_")
    try:
        os.chdir(tmp)
        with contextlib.redirect_stdout(io.StringIO()) as out:
            exec(compile(src, NE_PATH, "exec"), g)
        g["__stdout__"] = out.getvalue()
    finally:
        os.chdir(cwd)
        if not had_i64:
            del pd.Int64Index
        if not had_sa:
            del pd.Series.append
        if not had_da:
            del pd.DataFrame.append
    return g
