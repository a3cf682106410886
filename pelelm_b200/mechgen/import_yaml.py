"""Import a Cantera-YAML mechanism into the flat tables this package is built from.

Run HERE (where /root/reference exists); the resulting ``data/<mech>.json`` files are committed so
nothing reads the reference tree at run time.

Source of truth for the physics: ``Support/Mechanism/Models/<m>/mechanism.yaml``.
Conventions reproduced are those of the generated reference kernels in
``Support/Mechanism/Models/<m>/mechanism.H`` -- e.g. drm19 ``productionRate`` :3650-4902,
``aJacobian`` :9696-14742, thermo :342-1582 -- which is what the reference actually executes:

* reactions are ordered Troe falloff, Lindemann falloff, three-body, elementary
  (``ceptr/reaction_info.py:30-95``), original order inside each class;
* rate constants are in mol/m^3 units: A_SI = A * (1e-6)^(order-1), +1 order for "+ M";
* Ea/R uses Rc = 1.98721558317399615845 cal/mol/K (``ceptr/constants.py``);
* NASA-7 coefficients are rounded the way the generator printed them: 9 significant digits
  (``%+15.8e``) for cp/R, h/RT, d(cp/R)/dT and 16 (``%+20.15e``) for g/RT (``ceptr/thermo.py:357-432``);
* molecular weights are printed with 6 decimals (``mechanism.H:88-139``).

Numerical agreement of the tables with the compiled reference is asserted by
``tests/test_mech_tables_vs_ref.py`` (production rates, Jacobian, thermo on random states).
"""
from __future__ import annotations

import json
import os
import re
import sys

import yaml

RC_CAL = 1.98721558317399615845  # cal/mol/K
# Cantera 2.6 standard atomic weights (mechanism.cpp atomicWeight(); printed with %f)
ATOMIC_W = {"H": 1.008, "O": 15.999, "C": 12.011, "N": 14.007, "Ar": 39.95, "AR": 39.95, "He": 4.002602,
            "HE": 4.002602}

T_TROE, T_LIND, T_3B, T_SIMPLE = 3, 2, 1, 0


def _r(fmt: str, x: float) -> float:
    return float(fmt % x)


def _parse_side(side: str):
    """'H + 2 O2 + M' -> ({'H':1,'O2':2}, has_M)"""
    side = side.replace("(+M)", " ").replace("(+ M)", " ").replace("(+m)", " ")
    toks = [t.strip() for t in side.split(" + ")]
    out: dict[str, float] = {}
    has_m = False
    for t in toks:
        if not t:
            continue
        if t in ("M", "m"):
            has_m = True
            continue
        m = re.match(r"^(\d+(?:\.\d+)?)\s+(\S+)$", t)
        if m:
            coef, name = float(m.group(1)), m.group(2)
        else:
            coef, name = 1.0, t
        out[name] = out.get(name, 0.0) + coef
    return out, has_m


def import_mechanism(yaml_path: str, name: str) -> dict:
    with open(yaml_path) as f:
        doc = yaml.safe_load(f)
    units = doc.get("units", {})
    assert units.get("length") == "cm" and units.get("quantity") == "mol", units
    assert units.get("activation-energy") == "cal/mol", units
    species_names = list(doc["phases"][0]["species"])
    sp_by_name = {s["name"]: s for s in doc["species"]}
    N = len(species_names)
    idx = {s: i for i, s in enumerate(species_names)}

    mw, tmid, lo, hi = [], [], [], []
    for s in species_names:
        sp = sp_by_name[s]
        w = sum(ATOMIC_W[e] * n for e, n in sp["composition"].items())
        mw.append(float("%f" % w))
        th = sp["thermo"]
        assert th["model"] == "NASA7"
        tr = th["temperature-ranges"]
        if len(tr) == 3:
            tmid.append(float(tr[1]))
            lo.append([float(x) for x in th["data"][0]])
            hi.append([float(x) for x in th["data"][1]])
        else:  # single range: the generator sees mid = Tmax and identical halves
            tmid.append(float(tr[1]))
            lo.append([float(x) for x in th["data"][0]])
            hi.append([float(x) for x in th["data"][0]])

    def coef_tables(a):
        cp = [_r("%+15.8e", a[i]) for i in range(5)]
        h = [_r("%+15.8e", a[0]), _r("%+15.8e", a[1] / 2), _r("%+15.8e", a[2] / 3),
             _r("%+15.8e", a[3] / 4), _r("%+15.8e", a[4] / 5), _r("%+15.8e", a[5])]
        g = [_r("%+20.15e", a[5]), _r("%+20.15e", a[0] - a[6]), _r("%+20.15e", -a[0]),
             _r("%+20.15e", -a[1] / 2), _r("%+20.15e", -a[2] / 6), _r("%+20.15e", -a[3] / 12),
             _r("%+20.15e", -a[4] / 20)]
        dcp = [_r("%+15.8e", a[1]), _r("%+15.8e", a[2] * 2.0), _r("%+15.8e", a[3] * 3.0),
               _r("%+15.8e", a[4] * 4.0)]
        return cp, h, g, dcp

    thermo = {"tmid": tmid, "cp": [], "h": [], "g": [], "dcp": []}
    for k in range(N):
        tl, th_ = coef_tables(lo[k]), coef_tables(hi[k])
        for j, key in enumerate(("cp", "h", "g", "dcp")):
            thermo[key].append([tl[j], th_[j]])

    rxns = []
    for ri, r in enumerate(doc["reactions"]):
        eq = r["equation"]
        if " <=> " in eq:
            lhs, rhs = eq.split(" <=> ")
            rev = True
        elif " => " in eq:
            lhs, rhs = eq.split(" => ")
            rev = False
        else:
            lhs, rhs = eq.split(" = ")
            rev = True
        reac, m1 = _parse_side(lhs)
        prod, m2 = _parse_side(rhs)
        rtype = r.get("type", "elementary")
        for key in ("orders", "negative-A", "default-efficiency"):
            assert key not in r, (key, eq)
        order = sum(reac.values())
        d = {"orig": ri, "equation": eq, "reversible": rev,
             "reactants": [[idx[s], c] for s, c in reac.items()],
             "products": [[idx[s], c] for s, c in prod.items()],
             "eff": [[idx[s], float(e)] for s, e in r.get("efficiencies", {}).items()]}
        if rtype == "falloff":
            hp, lp = r["high-P-rate-constant"], r["low-P-rate-constant"]
            d["A"] = float(hp["A"]) * (1e-6) ** (order - 1)
            d["b"] = float(hp["b"])
            d["Ea_R"] = float(hp["Ea"]) / RC_CAL
            d["A0"] = float(lp["A"]) * (1e-6) ** order
            d["b0"] = float(lp["b"])
            d["Ea0_R"] = float(lp["Ea"]) / RC_CAL
            if "Troe" in r:
                t = r["Troe"]
                d["type"] = T_TROE
                d["troe"] = [float(t["A"]), float(t["T3"]), float(t["T1"]), float(t.get("T2", 0.0))]
                d["troe_has_T2"] = "T2" in t
            else:
                assert "SRI" not in r
                d["type"] = T_LIND
        elif rtype == "three-body":
            assert m1 and (m2 or not rev)
            assert len(d["eff"]) != 1, "single-collider three-body form is not handled: " + eq
            d["type"] = T_3B
            d["A"] = float(r["rate-constant"]["A"]) * (1e-6) ** order
            d["b"] = float(r["rate-constant"]["b"])
            d["Ea_R"] = float(r["rate-constant"]["Ea"]) / RC_CAL
        else:
            assert rtype == "elementary", rtype
            assert not m1
            d["type"] = T_SIMPLE
            d["A"] = float(r["rate-constant"]["A"]) * (1e-6) ** (order - 1)
            d["b"] = float(r["rate-constant"]["b"])
            d["Ea_R"] = float(r["rate-constant"]["Ea"]) / RC_CAL
        d["dn"] = sum(prod.values()) - sum(reac.values())
        rxns.append(d)

    ordered = ([r for r in rxns if r["type"] == T_TROE] + [r for r in rxns if r["type"] == T_LIND]
               + [r for r in rxns if r["type"] == T_3B] + [r for r in rxns if r["type"] == T_SIMPLE])
    return {"name": name, "species": species_names, "mw": mw, "thermo": thermo, "reactions": ordered,
            "n_species": N, "n_reactions": len(ordered)}


def main():
    ref = os.environ.get("PLM_REFERENCE", "/root/reference")
    models = os.path.join(ref, "Submodules/PelePhysics/Support/Mechanism/Models")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
    os.makedirs(out, exist_ok=True)
    for m in sys.argv[1:] or ["LiDryer", "drm19", "dodecane_lu"]:
        mech = import_mechanism(os.path.join(models, m, "mechanism.yaml"), m)
        with open(os.path.join(out, m + ".json"), "w") as f:
            json.dump(mech, f, indent=1)
        nr = [sum(1 for r in mech["reactions"] if r["type"] == t) for t in (3, 2, 1, 0)]
        print(m, "species", mech["n_species"], "reactions", mech["n_reactions"],
              "troe/lind/3b/simple", nr,
              "max slots", max(sum(c for _, c in r["reactants"]) for r in mech["reactions"]),
              max(sum(c for _, c in r["products"]) for r in mech["reactions"]),
              "irreversible", sum(1 for r in mech["reactions"] if not r["reversible"]))


if __name__ == "__main__":
    main()
