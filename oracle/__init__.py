"""TEST INFRASTRUCTURE ONLY -- nothing under upsp-processing_b200/ imports, links or executes anything in this package
(tests/test_abi.py checks that); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm do.

  upsp_oracle.c, upsp_oracle_setup.c, oracle.py   plain-C restatement of the psp_process frame chain and of the phase-0
                                                  projection matrix (every function cites the reference file:line it follows)
  ecc.py, setup_patches.py, p3d_overlap.py, targets.py   numpy / plain-Python restatements of the ECC solve, the patch
                                                  geometry, the structured model's seam detection and the target handling
  ref_probe.cpp + Makefile target `ref`           the REFERENCE'S OWN sources that compile in this image, built where they lie
                                                  under /root/reference into oracle/_ref/ (git-ignored): its two table tools, its
                                                  stand-alone transpose tool (mpi_stub/), and -- behind a small main() -- its deck,
                                                  paint-calibration, tunnel-condition and plot3d readers / writers, regression-sample
                                                  writer, peak finding (boost_stub/), unpack_12bit / unpack_10bit / MrawReader and
                                                  fix_hot_pixels (cv_stub/), its ray caster (BVH + watertight triangle test,
                                                  imath_stub/), its kd-tree and its patch geometry templates (eigen_stub/: an int mask matrix).
                                                  The stub headers carry no algorithm of the path.
How each restatement is pinned: DESIGN.md section 4.
"""
