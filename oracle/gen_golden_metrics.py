"""Generate tests/golden/prometheus_metrics_reference.json: name, kind, label names and histogram buckets of every metric the
reference service registers (presets/ragengine/metrics/prometheus_metrics.py, executed unmodified -- it only needs
prometheus_client).  kaito_b200/service.py must register the same set (tests/test_service.py).
Run: python oracle/gen_golden_metrics.py   (needs /root/reference)."""
import importlib.util
import json
import os

from prometheus_client.metrics import MetricWrapperBase

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "prometheus_metrics_reference.json")


def main():
    spec = importlib.util.spec_from_file_location("ref_metrics", "/root/reference/presets/ragengine/metrics/prometheus_metrics.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = {}
    for n in dir(m):
        v = getattr(m, n)
        if isinstance(v, MetricWrapperBase):
            ub = getattr(v, "_kwargs", {}).get("buckets")
            out[v._name] = {"kind": type(v).__name__, "labels": list(v._labelnames),
                            "buckets": None if type(v).__name__ != "Histogram" else [float(b) for b in (ub if ub is not None else v.DEFAULT_BUCKETS) if b != float("inf")]}
    consts = {k: getattr(m, k) for k in ("STATUS_SUCCESS", "STATUS_FAILURE", "MODE_LOCAL", "MODE_REMOTE")}
    json.dump({"meta": {"source": "presets/ragengine/metrics/prometheus_metrics.py executed unmodified", "generator": "oracle/gen_golden_metrics.py"},
               "constants": consts, "metrics": out}, open(OUT, "w"), indent=1, sort_keys=True)
    print(len(out), "metrics")


if __name__ == "__main__":
    main()
