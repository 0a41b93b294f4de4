"""Oracles for the SG-NN sparse hot path -- TEST INFRASTRUCTURE ONLY.

  O1  oracle/dense_equiv.py     dense conv3d identities (independent of any scn restatement)
  O2  oracle/sparseconvnet/     CPU restatement of the scn operator surface (drives the UNMODIFIED
                                /root/reference/torch/model.py; also the CPU timing baseline)
  O3  oracle/o3.c (+ o3.py)     fixed-summation-order C restatement, bit-exact target of the CUDA kernels

PARITY UNPINNED for the scn arithmetic: upstream SparseConvNet is not available offline and the reference
ships no tests/golden vectors.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
arms may import anything from this directory.
"""
