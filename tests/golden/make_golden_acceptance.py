"""Golden outputs of the reference's own acceptance drivers.

Runs the UNMODIFIED scripts of /root/reference as subprocesses on the reference's NumPy engine (this container
only -- the reference does not exist on the GPU box) and freezes their JSON under tests/golden/acceptance/.
tests/test_gpu_acceptance.py replays the same command lines through qsb.launcher on the B200 engine and compares.

    python tests/golden/make_golden_acceptance.py
"""

from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "acceptance")
REF = "/root/reference"

# name -> (script, argv); also imported by tests/test_gpu_acceptance.py so both sides run the same command lines
CASES = {
    # BASELINE config 4 exactly as SURVEY 8d states it (trials 100), all six metrics of all 15 points
    "qec_steane_depol_100": ("scripts/qec_threshold.py",
                             ["--codes", "steane", "--noise", "depolarizing", "--trials", "100", "--seed", "42"]),
    "qec_three_codes_bitflip_100": ("scripts/qec_threshold.py",
                                    ["--codes", "bit_flip,phase_flip,steane", "--noise", "bit_flip", "--trials", "100",
                                     "--seed", "7"]),
    "qec_two_codes_phaseflip_60": ("scripts/qec_threshold.py",
                                   ["--codes", "bit_flip,phase_flip", "--noise", "phase_flip", "--trials", "60",
                                    "--seed", "11"]),
    "noise_sweep_ghz3_depol": ("scripts/noise_sweep.py",
                               ["--circuit", "ghz3", "--noise", "depolarizing", "--steps", "4", "--trials", "20",
                                "--seed", "42"]),
    "noise_sweep_bell_bitflip": ("scripts/noise_sweep.py",
                                 ["--circuit", "bell", "--noise", "bit_flip", "--steps", "6", "--trials", "30",
                                  "--seed", "5"]),
    "vqe_heisenberg_3q": ("scripts/vqe_benchmark.py",
                          ["--qubits", "3", "--layers", "2", "--hamiltonian", "heisenberg", "--iters", "5",
                           "--seed", "42"]),
    "vqe_z0_2q": ("scripts/vqe_benchmark.py", ["--qubits", "2", "--layers", "2", "--iters", "8", "--seed", "3"]),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (script, argv) in CASES.items():
        with tempfile.NamedTemporaryFile(suffix=".json", delete=False) as f:
            tmp = f.name
        cmd = [sys.executable, os.path.join(REF, script)] + argv + ["--output", tmp]
        subprocess.run(cmd, check=True, cwd=REF, stdout=subprocess.DEVNULL)
        with open(tmp) as f:
            data = json.load(f)
        os.unlink(tmp)
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump({"script": script, "argv": argv, "output": data}, f, indent=1, sort_keys=True)
        print("wrote", name)
    # test_validation.py: the transcript of the real reference (33/33) -- the replay must print the same PASS lines
    res = subprocess.run([sys.executable, os.path.join(REF, "test_validation.py")], cwd=REF, capture_output=True,
                         text=True, check=True)
    lines = [ln.strip() for ln in res.stdout.splitlines() if ln.strip().startswith("[") or ln.startswith("Results:")]
    with open(os.path.join(OUT, "test_validation_transcript.json"), "w") as f:
        json.dump({"lines": lines}, f, indent=1)
    print("wrote test_validation transcript:", lines[-1])


if __name__ == "__main__":
    main()
