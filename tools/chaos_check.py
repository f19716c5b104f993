"""Protocol test of the tensor-core kernels: run with AFT_B200_LIB pointing at the "chaos" build (random delays
before every mbarrier wait).  The bf16 forward must still match the fp32 path (which does not use the protocol) and no
wait may time out.  Prints one JSON line; exit code 0 on success."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adafortitran_b200 import _capi
from oracle import aft_oracle as O          # synthetic inputs only
from tests import util

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
    kind = sys.argv[2] if len(sys.argv) > 2 else "forti"
    sd = util.ada_weights()
    if kind != "ada": sd = util.forti_weights(sd)
    p, snr, ds, dop = O.synthetic_batch(B, seed=5)
    md = util.meta(snr, ds, dop) if kind == "ada" else None
    torch.zeros(1, device="cuda")
    err = C.c_double(-1)
    _capi.lib().aft_selftest(100, C.byref(err), None)          # arm the wait-timeout diagnostics
    out = {}
    failure = None
    try:
        for prec in ("fp32", "bf16"):
            m = util.make_model(kind, weights=sd, precision=prec)
            with torch.no_grad():
                out[prec] = m(torch.from_numpy(p), md).cpu().numpy()
        torch.cuda.synchronize()
    except Exception as e:                                      # a protocol hang surfaces as a trapped kernel
        failure = str(e)[:200]
    _capi.lib().aft_selftest(101, C.byref(err), None)
    res = {"lib": os.path.basename(_capi.lib_path()), "B": B, "kind": kind, "wait_timeouts": err.value, "failure": failure}
    if failure is None:
        res["rel_db_bf16_vs_fp32"] = O.rel_err_db(out["bf16"], out["fp32"])
        res["finite"] = bool(np.isfinite(out["bf16"].view(np.float32)).all())
    print(json.dumps(res), flush=True)
    ok = failure is None and err.value == 0 and res["finite"] and res["rel_db_bf16_vs_fp32"] <= (-45.0 if kind != "ada" else -36.0)
    sys.exit(0 if ok else 1)

if __name__ == "__main__":
    main()
