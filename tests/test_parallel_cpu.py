"""N > 1 host logic on CPU: two `gloo` ranks each own the pixels yh_shard_pixels
gives them, fill ONLY those from a render, and the single reduce at the end must
reproduce the whole image bit for bit.  The CPU checker stands in
for the device renderer here (this is a test of sharding + reduce, not of kernels)."""
import os
import socket
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, {pkg!r}); sys.path.insert(0, {orc!r})
    import oracle
    from yocto_hair_b200 import parallel, synth
    from yocto_hair_b200.scene import make_params
    rank, local, world = parallel.env_rank()
    dist.init_process_group("gloo")
    assert dist.get_world_size() == world == 2
    sc, cam = synth.curly_hair_scene(1200)
    o = oracle.best_oracle()
    rs = o.scene(sc); st = rs.state(cam, make_params(resolution=96)); st.trace_samples(2)
    _, accum, samples, _ = st.download()
    mask = parallel.shard_mask(st.width, st.height, rank, world)
    mine_a = torch.from_numpy(np.where(mask[..., None], accum, 0).astype(np.float32))
    mine_s = torch.from_numpy(np.where(mask, samples, 0).astype(np.int32))
    counts = torch.tensor([int(mask.sum())]); dist.all_reduce(counts)
    assert counts.item() == st.width * st.height
    kept_a = mine_a.clone()
    mine_a_in, mine_s_in = mine_a, mine_s
    mine_a, mine_s = parallel.reduce_image(mine_a_in, mine_s_in, 0)
    assert torch.equal(mine_a_in, kept_a)  # the inputs are left untouched: the sum goes to scratch
    if rank == 0:
        assert np.array_equal(mine_s.numpy(), samples)
        assert np.array_equal(mine_a.numpy().view(np.uint32), accum.view(np.uint32))
        r = parallel.resolve(mine_a, mine_s).numpy()
        assert np.array_equal(r.view(np.uint32), (accum / samples[..., None].astype(np.float32)).view(np.uint32))
        print("PARALLEL_OK", int(mask.sum()))
    dist.barrier(); dist.destroy_process_group()
""")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_shard_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(pkg=str(ROOT / "yocto-hair_b200"), orc=str(ROOT / "oracle")))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1", GLOO_SOCKET_IFNAME="lo")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "PARALLEL_OK" in out.stdout


def test_bench_reference_arm_other_ranks_exit_quietly(tmp_path):
    """`bench.py --impl reference` under torchrun: only rank 0 runs and prints."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "tiny"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
