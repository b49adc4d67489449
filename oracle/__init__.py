"""CPU oracle for the pair-matching hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (reconstructor_b200/) never does.

  oracle.orc      ctypes binding of oracle/libpm_oracle.so (pm_oracle.c, plain C)
  oracle.cv2_ref  the reference's per-pair body restated over Python cv2 (the library
                  that carries the reference's arithmetic); used to pin pm_oracle.c and
                  to generate tests/golden/, and as the timed CPU arm of bench.py.
"""
