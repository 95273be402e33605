"""Per-chunk timeline of one hpf_solve_host call ($HPF_HOST_TIMELINE) for a list of chunk plans."""
import os, sys, tempfile
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch, helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
net, _, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = 65536
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight")
hP, hQ, hI = (torch.as_tensor(x).pin_memory().numpy() for x in (P, Q, I_N))
for plan in sys.argv[1:]:
    os.environ["HPF_HOST_PLAN"] = plan
    os.environ.pop("HPF_HOST_TIMELINE", None)
    for _ in range(3):
        sol.solve_host(hP, hQ, hI)
    os.environ["HPF_HOST_TIMELINE"] = "1"
    sys.stderr.write("== plan %s\n" % plan); sys.stderr.flush()
    sol.solve_host(hP, hQ, hI)
