"""Derive the object -> target-class lookup tables from the reference's annotation JSONs.

Runs only in the builder container (needs /root/reference); the output hoigen_b200/data/object_tables.json
(small integer lists) is committed.  Derivations restate, not copy:
  - hico_object_to_verb / hico_object_to_interaction / object_n_verb_to_interaction:
        hicodet/hicodet.py:145-185 over instances_test2015.json['correspondence'] = [hoi, obj, verb] x 600
  - vcoco_object_to_action: vcoco/vcoco.py:153-160 over instances_vcoco_test.json['action_to_object']
        then `list(object_to_action.values())` as main_tip_finetune.py:848 does (objects 1..80 -> rows 0..79).
  - hico_unseen_uc0: hico_text_label.py:827-840 `hico_unseen_index['uc0']`, the 120 HOI ids the UC zero-shot split
        (`--zs --zs_type uc0`, BASELINE configs[2]) holds out.
"""
import json
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "hoigen_b200" / "data" / "object_tables.json"


def main():
    corr = json.load(open(REF / "hicodet" / "instances_test2015.json"))["correspondence"]
    assert len(corr) == 600
    o2v = [[] for _ in range(80)]
    o2i = [[] for _ in range(80)]
    onv = [[-1] * 117 for _ in range(80)]
    for hoi, obj, verb in corr:
        o2v[obj].append(verb)
        o2i[obj].append(hoi)
        onv[obj][verb] = hoi
    a2o = json.load(open(REF / "vcoco" / "instances_vcoco_test.json"))["action_to_object"]
    o2a = {o: [] for o in range(1, 81)}
    for act, objs in enumerate(a2o):
        for o in objs:
            if act not in o2a[o]:
                o2a[o].append(act)
    ns = {}
    src = (REF / "hico_text_label.py").read_text()
    exec(compile(src[src.index("hico_unseen_index = {"):].split("\n}\n")[0] + "\n}\n", "hico_unseen_index", "exec"), ns)
    uc0 = [int(v) for v in ns["hico_unseen_index"]["uc0"]]
    assert len(uc0) == 120 and len(set(uc0)) == 120 and max(uc0) < 600
    tables = dict(
        hico_object_to_verb=o2v,
        hico_object_to_interaction=o2i,
        hico_object_n_verb_to_interaction=onv,   # -1 where the reference has None
        hico_correspondence=corr,
        vcoco_object_to_action=list(o2a.values()),
        hico_unseen_uc0=uc0,
    )
    OUT.parent.mkdir(exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(tables, f, separators=(",", ":"))
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
