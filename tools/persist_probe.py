"""Per-step cost of the persistent step kernel against the CUDA-graph path, with the rebuild cadence varied (what a rebuild cycle
costs on top of its plain steps).  python tools/persist_probe.py [workload]   (GPU box)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(workload="ab_gas"):
    from pfmds_b200 import inputs
    from pfmds_b200.engine import configure
    for period in (20, 100, 400):
        if workload == "ab_gas":
            case, integ, dt = inputs.ab_gas(period=period), "nvt", 0.5
        else:
            case, integ, dt = inputs.graphene_on_cu(period=period), "nvt", 1.0
        for env in ({"PFMDS_PERSIST": "0"}, {}, {"PFMDS_PERSIST_BLOCKS_PER_SM": "1"}, {"PFMDS_PERSIST_BLOCKS_PER_SM": "2"}):
            saved = {k: os.environ.get(k) for k in env}
            os.environ.update(env)
            try:
                eng = configure(case)
            finally:
                for k, v in saved.items():
                    os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
            eng.advance(integ, dt, 0, 41)
            eng.synchronize()
            l0 = eng.launch_count()
            eng.timer_start()
            eng.advance(integ, dt, 41, 1600)
            ms = eng.timer_stop()
            print("%s period %3d %-36s %.2f us/step  launches/step %.2f" % (workload, period, env or "default", ms / 1600 * 1e3, (eng.launch_count() - l0) / 1600.0), flush=True)
            eng.close()


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["ab_gas"]))
