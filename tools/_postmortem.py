"""Decoder of the train-mode tower's post-mortem words (csrc/mnk_resnet_train.cu: written by the commit watcher of the first CTA
that times out; sticky until the scratch buffer is zeroed).  Words 14.. only exist in a -DMNK_PROGRESS build."""
import torch

STAGES = {"epi": {1: "loop top", 2: "produced, before token sync", 3: "token passed", 4: "TMEM loaded", 5: "stored, before release",
                  6: "released", 7: "left the loop"},
          "mma": {1: "before step-barrier sync", 2: "step barrier passed", 3: "issued + committed", 7: "left the loop"},
          "watch": {1: "waiting for the commit", 2: "token sent"}}


def read(native, lib, m, n, envs):
    need = int(lib.mnk_resnet_tower_train_scratch_bytes(m, n, envs, native.blocks))
    return native._scratch[need - 256 + 64: need - 256 + 256].view(torch.int32).tolist()


def describe(pm, m=9):
    u = [w & 0xFFFFFFFF for w in pm]
    tot = (u[1] >> 16) & 0xFFFF
    e = u[2] & 0xFFFF
    out = [f"    first watcher timeout since the scratch was zeroed ({u[0]} in all): cta {u[1] & 0xFFFF}, {u[3]} groups, {tot} steps, "
           f"watcher at step {e} (total-{tot - e}, row {e % m}), layer {u[2] >> 16}",
           "    commit barriers " + " ".join(hex((u[5 + 2 * q] << 32) | u[4 + 2 * q]) for q in range(5))]
    if any(u[14:32]):
        for w in range(18):
            role = "epi" if w < 16 else ("mma" if w == 16 else "watch")
            idx, st = u[14 + w] >> 4, u[14 + w] & 15
            out.append(f"      warp {w:2d} ({role}{'' if w >= 16 else ' set %d' % (w // 8)}): index {idx} (total-{tot - idx}), "
                       f"{STAGES[role].get(st, st)}")
        out.append("    operand barriers " + " ".join(hex((u[33 + 2 * q] << 32) | u[32 + 2 * q]) for q in range(2)))
    return "\n".join(out)
