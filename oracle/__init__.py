"""oracle/ — CPU restatement of the reference's algorithms for the Drone-YOLO inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under drone_yolo_b200/ imports this package; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs may use it, and only as the checker / the CPU baseline.

Parity status: PINNED.  The reference ships no tests or golden vectors for this path, so the
restatement is pinned against outputs of the reference itself, generated in the authoring container by importing
/root/reference (tools/make_golden.py -> tests/golden/*.npz) and re-checked by tests/test_oracle_golden.py.
torchvision.ops.nms (torchvision 0.26.0, un-vendored third-party dependency, call site ultralytics/utils/ops.py:312)
is restated in nms_np.nms_greedy and cross-checked against the installed torchvision in the same tests.
"""
