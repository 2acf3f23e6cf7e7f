"""Discrete-event model of one backward CTA's steady state (attn_bwd_sm100.cu): which dependency cycle bounds a step,
and what a different TMEM aliasing / stage count would buy.  Durations are the measured ones (tools/mma_mix_bench.cu,
clock64 traces under profiles/).  usage: python tools/bwd_pipeline_sim.py"""
import sys

def simulate(scheme, NQ=2, NDO=2, n=40, tma=1900, hop=150, mma_scale=1.0, a_ld=250, a_rest=1150, b_ld=200, b_mid=150, b_st=150,
             drain_ld=200):
    D = {k: v * mma_scale for k, v in dict(S=490, dP=490, dV=640, dQ=560, dK=560).items()}
    INF = float("inf")
    # event times
    ev = {}
    def get(k):
        return ev.get(k, None)
    # we iterate to a fixed point by processing ops in a global loop until all defined
    # streams: lists of (op, i, deps) in order
    def deps_of(op, i):
        d = []
        if scheme == "old":
            if op == "S":  d = [("Qfull", i)] + ([("Pready", i - 1)] if i > 0 else [])
            if op == "dP": d = [("dOfull", i), ("Qfull", i)] + ([("DQdrained", i - 1)] if i > 0 else [])
            if op == "dV": d = [("Pready", i), ("dOfull", i)]
            if op == "dQ": d = [("DSready", i)]
            if op == "dK": d = [("DSready", i)]
        else:
            if op == "S":  d = [("Qfull", i)] + ([("Pready", i - 1)] if i > 0 else []) + ([("DQdrained", i - 2)] if i > 1 else [])
            if op == "S" and scheme in ("new1", "new1s"): d = [("Qfull", i)] + ([("Pready", i - 1)] if i > 0 else []) + ([("DQdrained", i - 2)] if i > 1 else [])
            if op == "dP": d = [("dOfull", i), ("Qfull", i)] + ([("DPtaken", i - 1)] if i > 0 else [])
            if op == "dV": d = [("Pready", i), ("dOfull", i)]
            if op == "dQ": d = [("DSready", i)] + ([("Staken", i + 1)] if i + 1 < n else [])
            if op == "dK": d = [("DSready", i)]
        return d
    if scheme == "old":
        A = [("S", 0)]; B = [("dP", 0)]
        for i in range(n):
            A += [("dV", i)] + ([("S", i + 1)] if i + 1 < n else [])
            B += [("dQ", i), ("dK", i)] + ([("dP", i + 1)] if i + 1 < n else [])
    elif scheme in ("new1", "new1s"):
        A = [("S", 0)]; B = [("dP", 0)]
        for i in range(n):
            if scheme == "new1": A += [("dV", i)] + ([("S", i + 1)] if i + 1 < n else [])
            else: A += ([("S", i + 1)] if i + 1 < n else []) + [("dV", i)]
            B += ([("dP", i + 1)] if i + 1 < n else []) + [("dK", i), ("dQ", i)]
    else:
        A = [("S", 0)] + ([("S", 1)] if n > 1 else []); B = [("dP", 0)]
        for i in range(n):
            A += [("dV", i)] + ([("S", i + 2)] if i + 2 < n else [])
            B += ([("dP", i + 1)] if i + 1 < n else []) + [("dK", i), ("dQ", i)]
    if scheme == "new" :
        pass
    # pointers
    pa = pb = 0
    issue = {}      # (op,i) -> issue time (when the issuing thread got past its waits)
    done = {}       # (op,i) -> completion time in the pipe
    pipe_free = 0.0
    thread_free = {"A": 0.0, "B": 0.0}
    # compute / drain / tma state
    comp_t = 0.0; comp_seq = []
    for i in range(n):
        comp_seq += [("a", i), ("b", i)]
    pc = 0
    drain_t = 0.0; pd = 0
    qload = {}; doload = {}
    prod_t = 0.0; pq = 0; pdo = 0
    def ready_time(deps):
        t = 0.0
        for d in deps:
            if d not in ev: return None
            t = max(t, ev[d] + hop)
        return t
    progress = True
    pending = []   # issued, not yet executed in pipe: (issue_time, stream, op, i)
    while progress:
        progress = False
        # TMA producer: Q(i) then dO(i) alternately, each needs its stage free
        while pq < n or pdo < n:
            moved = False
            if pq < n and pq <= pdo:
                need = ("Qfree", pq - NQ) if pq >= NQ else None
                if need is None or need in ev:
                    t = max(prod_t, (ev[need] + hop) if need else 0.0)
                    ev[("Qfull", pq)] = t + tma; prod_t = t + 20; pq += 1; moved = True
            elif pdo < n:
                need = ("dOfree", pdo - NDO) if pdo >= NDO else None
                if need is None or need in ev:
                    t = max(prod_t, (ev[need] + hop) if need else 0.0)
                    ev[("dOfull", pdo)] = t + tma; prod_t = t + 20; pdo += 1; moved = True
            if not moved: break
            progress = True
        # issuing threads
        for name, lst in (("A", A), ("B", B)):
            p = pa if name == "A" else pb
            while p < len(lst):
                op, i = lst[p]
                rt = ready_time(deps_of(op, i))
                if rt is None: break
                t = max(rt, thread_free[name])
                issue[(op, i)] = t; thread_free[name] = t + 40
                pending.append((t, name, op, i)); p += 1; progress = True
            if name == "A": pa = p
            else: pb = p
        # pipe executes in issue order
        pending.sort()
        while pending:
            t, name, op, i = pending[0]
            # an op can only run if every earlier-issued op has run (FIFO) -- pending sorted by issue time; but a later
            # fixed-point round may add an earlier one; accept the approximation (streams issue monotonically)
            start = max(t, pipe_free)
            end = start + D[op]
            pipe_free = end; done[(op, i)] = end; pending.pop(0); progress = True
            if op == "S": ev[("Sfull", i)] = end
            if op == "dP": ev[("DPfull", i)] = end
            if op == "dQ": ev[("DQfull", i)] = end
            if op in ("S", "dK") and ("S", i) in done and ("dK", i) in done:
                ev[("Qfree", i)] = max(done[("S", i)], done[("dK", i)])
            if op in ("dP", "dV") and ("dP", i) in done and ("dV", i) in done:
                ev[("dOfree", i)] = max(done[("dP", i)], done[("dV", i)])
        # compute warps
        while pc < len(comp_seq):
            ph, i = comp_seq[pc]
            if ph == "a":
                if ("Sfull", i) not in ev: break
                t = max(comp_t, ev[("Sfull", i)] + hop) + a_ld
                ev[("Staken", i)] = t
                t += a_rest
                ev[("Pready", i)] = t; comp_t = t
            else:
                if ("DPfull", i) not in ev: break
                t = max(comp_t, ev[("DPfull", i)] + hop) + b_ld
                ev[("DPtaken", i)] = t
                t += b_mid
                if scheme != "old" and i > 0:
                    if ("DQfull", i - 1) not in ev: break
                    t = max(t, ev[("DQfull", i - 1)] + hop)
                t += b_st
                ev[("DSready", i)] = t; comp_t = t
            pc += 1; progress = True
        # drain
        while pd < n:
            if ("DQfull", pd) not in ev: break
            t = max(drain_t, ev[("DQfull", pd)] + hop) + drain_ld
            ev[("DQdrained", pd)] = t
            drain_t = t + 900     # staging + reductions of the three atoms
            pd += 1; progress = True
    ds = [ev[("DSready", i)] for i in range(n) if ("DSready", i) in ev]
    if len(ds) < n: return None
    return (ds[-1] - ds[n // 2]) / (n - 1 - n // 2)

if __name__ == "__main__":
    for tma in (1900, 1200):
        for scale in (1.0, 0.8):
            print(f"TMA latency {tma}, MMA cost x{scale}")
            for scheme, NQ, NDO in (("old", 2, 2), ("new1", 2, 2), ("new1", 3, 2), ("new1s", 2, 2), ("new1s", 3, 2), ("new1s", 3, 3), ("new", 3, 2), ("new", 3, 3)):
                r = simulate(scheme, NQ, NDO, tma=tma, mma_scale=scale)
                print(f"   {scheme:4s} Q stages {NQ} dO stages {NDO}: {r and round(r)} cycles per step")
