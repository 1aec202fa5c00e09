"""What does the hand-over cost a peer?  Under torchrun (N >= 2): the device-timed loop of bench.py (p2p gather) with parts
of the protocol switched off — no frame gate, no delivery signal (and no wait on rank 0), stores kept local.  Prints every
rank's [gate + draw, whole step, draw alone] in microseconds per variant.  Timing experiment only: frames are not checked."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class P(bench.Pass):
    mode = "full"

    def render_only(self, rot=None, light4=None):
        self.fno += 1
        f = self.fno
        dev_ptr, delivered, consumed = self.slot_ptrs(f)
        self.uses[f & 1] += 1
        if self.rank != 0:
            if "nogate" not in self.mode:
                self.r.gate_next_frame(consumed, max(f - 2, 0))
            if "nosignal" not in self.mode:
                self.r.signal_after_frame(delivered)
            elif "memop" in self.mode:
                pass
        self.r.render_device(self.rot, self.cam4, self.light4, self.cfg.focal, dev_ptr=(0 if "local" in self.mode else dev_ptr), stream=self.sptr)
        if self.rank != 0 and "memop" in self.mode:  # per-peer word written by a stream memory operation instead of a kernel
            self.r.stream_write(delivered + 4 * (8 + self.rank), self.uses[f & 1], self.sptr)

    def gather_only(self, consume=True):
        if self.rank == 0:
            _, delivered, consumed = self.slot_ptrs(self.fno)
            if "memop" in self.mode:
                for k in range(1, self.world):
                    self.r.stream_wait_geq(delivered + 4 * (8 + k), self.uses[self.fno & 1], self.sptr)
            elif "nosignal" not in self.mode:
                self.r.stream_wait_geq(delivered, (self.world - 1) * self.uses[self.fno & 1], self.sptr)
            self.r.stream_write(consumed, self.fno, self.sptr)


def main():
    import torch
    import torch.distributed as dist
    import uob_raytracer_b200 as u
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    args = argparse.Namespace(no_parallel_egress=True, steps=100, warmup=5)
    cfg = u.CONFIGS["cfg2"]
    out = {}
    modes = [a for a in sys.argv[1:]] or ["full", "nogate", "nosignal", "nosignal+memop", "nogate+nosignal", "nogate+nosignal+local", "full"]
    for mode in modes:
        p = P(args, cfg, world, rank, local, dist, "p2p", False)
        p.mode = mode
        t = p.time_device(100, 5, False)
        out.setdefault(mode + ("" if mode not in out else " "), []).append({"step_us": round(t["ms_per_step"] * 1e3, 1), "per_rank_us": [[round(x * 1e3, 1) for x in r] for r in t["per_rank"]]})
        p.close()
        dist.barrier()
    if rank == 0:
        for k, v in out.items():
            for q in v:
                print(f"{k:24s} step {q['step_us']:6.1f} us   per rank [gate+draw, step, draw alone]: {q['per_rank_us']}")
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"peer_cost_n{world}.json"), "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
