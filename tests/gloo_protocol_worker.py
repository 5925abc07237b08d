"""One rank of the world_size-2 CPU test (gloo) of gen.run_distributed's protocol: the engine is a stub,
what is tested is that the ranks agree on the outcome -- everybody returns, everybody starts over
(a streamed plan's bounds did not hold), or everybody raises when ONE rank fails."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(rank: int, world: int, port: int, out_path: str):
    import torch.distributed as dist
    import genlib_b200 as gen
    engine_mod = sys.modules[gen.run_distributed.__module__]

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakePlan:
        def __init__(self):
            self.world = world

    script = {}                                   # attempt -> what this rank's run() does

    class StubEngine:
        created = 0

        def __init__(self, plan, numerics="reference", device=0, rank=None):
            self.attempt = StubEngine.created
            StubEngine.created += 1
            self.closed = False

        def ipc_handle(self):
            return bytes(64)

        def attach(self, handles):
            assert len(handles) == world

        def run(self):
            what = script.get(self.attempt, "ok")
            if what == "restart":
                raise gen.PlanBoundsExceeded(7, "bounds")
            if what == "fail":
                raise gen.GenlibError(4, "boom")

        def close(self):
            self.closed = True

    engine_mod.Engine = StubEngine
    results = []
    # 1. everybody fine
    StubEngine.created = 0; script.clear()
    e = gen.run_distributed(FakePlan(), rank=rank)
    results.append(e.attempt == 0 and not e.closed)
    # 2. every rank's first run asks for a restart (the same plan, the same bounds): second engine runs
    StubEngine.created = 0; script.clear(); script[0] = "restart"
    e = gen.run_distributed(FakePlan(), rank=rank)
    results.append(e.attempt == 1 and not e.closed)
    # 3. ONE rank fails: both raise, nobody is left waiting
    StubEngine.created = 0; script.clear()
    if rank == 1:
        script[0] = "fail"
    try:
        gen.run_distributed(FakePlan(), rank=rank)
        results.append(False)
    except gen.GenlibError as err:
        results.append(rank == 1 and "boom" in str(err))
    except RuntimeError as err:
        results.append(rank == 0 and "rank(s) [1] failed" in str(err))
    # 4. the restart does not help either: both raise PlanBoundsExceeded
    StubEngine.created = 0; script.clear(); script[0] = "restart"; script[1] = "restart"
    try:
        gen.run_distributed(FakePlan(), rank=rank)
        results.append(False)
    except gen.PlanBoundsExceeded:
        results.append(True)
    flags = [None] * world
    dist.all_gather_object(flags, all(results))
    dist.destroy_process_group()
    if rank == 0:
        with open(out_path, "w") as fh:
            fh.write("ok" if all(flags) else "fail " + repr(results))


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
