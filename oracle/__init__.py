"""CPU oracle for the DORAEMON hot path — TEST INFRASTRUCTURE ONLY.

Plain numpy / PyTorch-fp32(/fp64) restatements of the reference's algorithms, each function citing the
reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (visiondk_b200/) never does and fails loudly
when its CUDA library is missing.

Pinning status:
  * heads / CE / EMA / SGD / scheduler: PINNED — checked against the reference's own modules executed in
    the authoring container (oracle/make_golden.py imports /root/reference files by path; outputs are
    committed under tests/golden/).
  * ConvNeXt backbone (timm 0.9.16, not vendored in the reference) and flat inner-product search
    (faiss-gpu 1.8.0, not vendored): PARITY UNPINNED — the reference holds no tests or golden vectors at
    those boundaries and neither package is installed; the restatements follow the published algorithms
    and are cross-checked against torchvision's architecture-identical convnext_base and against a direct
    fp64 evaluation respectively.
"""
