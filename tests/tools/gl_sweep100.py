import sys, os, json
sys.path.insert(0, '/root/repo/tests/tools'); sys.path.insert(0, '/root/repo')
from gl_modes_check import *
rows = []
kinds = [("clip%d" % c, synth_clip(c, 10.0)) for c in range(4)]
kinds.append(("white_noise", np.random.default_rng(77).uniform(-1, 1, 441000)))
kinds.append(("silence", np.zeros(441000)))
leads = (8, 16, 32, 50, 72, 90)
for name, wav in kinds:
    mel = O.to_mel(O.config(), wav)
    for s in range(16):
        seed = 100 + s
        init = np.random.default_rng(seed).random(440576)
        f64 = run(mel, init, 100, True)
        row = {"clip": name, "seed": seed}
        for lead in leads:
            row["lead%d" % lead] = rel_l2(run(mel, init, 100, False, lead=lead), f64)
        rows.append(row)
    sub = [r for r in rows if r["clip"] == name]
    print(name, " ".join("lead%d max %.1e" % (l, max(r["lead%d" % l] for r in sub)) for l in leads), flush=True)
# trace of two bad cases
for name, seed in (("clip2", 111), ("white_noise", 102)):
    wav = dict(kinds)[name]; mel = O.to_mel(O.config(), wav)
    init = np.random.default_rng(seed).random(440576)
    line = f"trace {name}/{seed} lead8:"
    for iters in (10, 20, 30, 40, 50, 60, 70, 80, 90, 100):
        line += f" {iters}:{rel_l2(run(mel, init, iters, False, lead=8), run(mel, init, iters, True)):.1e}"
    print(line, flush=True)
json.dump(rows, open('/root/repo/gpurun_out/gl_parity_sweep100.json', 'w'), indent=1)
