"""The C++ driver (bin/miniVite_b200) on one GPU: reference option surface + report lines + parity."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bin", "miniVite_b200")


def test_cli_generate_matches_reference(tmp_path, golden):
    case = golden["rgg_n16384_p1"]
    p = subprocess.run([EXE, "-n", "16384", "-T", "-o", str(tmp_path / "c")], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert f"Number of edges: {case['ne']}" in p.stdout
    assert "64-bit datatype" in p.stdout and "Average total time (in s), #Processes: " in p.stdout
    m = re.search(r"Modularity, #Iterations: (\S+), (\d+)", p.stdout)
    assert int(m.group(2)) == case["iters"] and abs(float(m.group(1)) - float(case["modularity"])) < 1e-5
    it = re.findall(r"ITER (\d+) mod=(\S+) moved=(\d+) chash=([0-9a-f]+)", p.stderr)
    assert [(float(a[1]), int(a[2]), a[3]) for a in it] == [(float(g["modularity"]), g["moved"], g["chash"]) for g in case["trace"]]
    raw = np.fromfile(str(tmp_path / "c.0"), dtype=np.int64)
    assert raw[0] == 0 and raw[1] == 16384
    from oracle import oracle as O
    assert "%016x" % O.comm_hash(0, raw[2:]) == case["final_chash"]


def test_cli_file_and_weighted(tmp_path, golden):
    from minivite_b200 import hostgraph as hg
    ss = hg.generate_rgg(16384, 4)
    path = str(tmp_path / "g.bin")
    ss.write(path)
    p = subprocess.run([EXE, "-f", path, "-T"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    case = golden["file_rgg_n16384_s4_p1"]
    it = re.findall(r"ITER (\d+) mod=(\S+) moved=(\d+) chash=([0-9a-f]+)", p.stderr)
    assert [(float(a[1]), int(a[2]), a[3]) for a in it] == [(float(g["modularity"]), g["moved"], g["chash"]) for g in case["trace"]]
    p = subprocess.run([EXE, "-n", "16384", "-w"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0
    m = re.search(r"RESULT mod=(\S+) iters=(\d+)", p.stderr)
    assert abs(float(m.group(1)) - float(golden["rgg_n16384_p1_w"]["modularity"])) <= 1e-6


def test_cli_rejects_bad_options():
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode != 0 and "Must specify some options." in p.stderr
    p = subprocess.run([EXE, "-n", "1000", "-g", "3"], capture_output=True, text=True)
    assert p.returncode != 0


def test_cli_device_generation(tmp_path, golden):
    """-D: GenerateRGG runs on the GPU (mvgpu_generate_rgg_shard); same graph, same trace as the reference's `-n`."""
    case = golden["rgg_n16384_p1"]
    p = subprocess.run([EXE, "-n", "16384", "-T", "-D"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    assert f"Number of edges: {case['ne']}" in p.stdout
    it = re.findall(r"ITER (\d+) mod=(\S+) moved=(\d+) chash=([0-9a-f]+)", p.stderr)
    assert [(float(a[1]), int(a[2]), a[3]) for a in it] == [(float(g["modularity"]), g["moved"], g["chash"]) for g in case["trace"]]


def test_cli_ranks_sharing_one_device(golden):
    """`-g 2` on a one-GPU box: the ranks (forked processes, like `mpirun -n 2`) wrap around onto device 0 and bootstrap
    through the host transport.  Host-generated and device-generated (-D) graphs, default RNG and -l: traces equal the
    reference's own 2-rank runs of `miniVite -n 16384 [-l]` (goldens from the reference's generator)."""
    import os
    env = dict(os.environ, MVGPU_OPTIONS="host_transport=1")
    for args, name in ((["-n", "16384"], "rgg_n16384_p2"), (["-n", "16384", "-D"], "rgg_n16384_p2"),
                       (["-n", "16384", "-l"], "rgg_n16384_p2_l"), (["-n", "16384", "-l", "-D"], "rgg_n16384_p2_l")):
        case = golden[name]
        p = subprocess.run([EXE, "-g", "2", "-T"] + args, capture_output=True, text=True, timeout=300, env=env)
        assert p.returncode == 0, p.stderr[-2000:]
        assert f"Number of edges: {case['ne']}" in p.stdout
        it = re.findall(r"ITER (\d+) mod=(\S+) moved=(\d+) chash=([0-9a-f]+)", p.stderr)
        assert [(float(a[1]), int(a[2]), a[3]) for a in it] == [(float(g["modularity"]), g["moved"], g["chash"]) for g in case["trace"]], args


def test_cli_device_generation_with_random_edges():
    """-D -p: the random long edges are generated on the GPU too; report lines and trace equal the host-generated run
    (single rank, and two ranks sharing device 0 where the generator sums the edge count over the ranks)."""
    env = dict(os.environ, MVGPU_OPTIONS="host_transport=1")
    for extra in ([], ["-g", "2"]):
        outs = []
        for dev in ([], ["-D"]):
            p = subprocess.run([EXE, "-n", "32768", "-p", "10", "-T"] + extra + dev, capture_output=True, text=True, timeout=300, env=env)
            assert p.returncode == 0, p.stderr[-2000:]
            ne = re.search(r"Number of edges: (\d+)", p.stdout).group(1)
            mod = re.search(r"Modularity, #Iterations: (\S+), (\d+)", p.stdout).groups()
            it = re.findall(r"ITER (\d+) mod=(\S+) moved=(\d+) chash=([0-9a-f]+)", p.stderr)
            assert len(it) == int(mod[1])
            outs.append((ne, mod, it))
        assert outs[0] == outs[1], extra
