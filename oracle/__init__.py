"""CPU oracle for golfer-b200 — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (/root/reference) ships a README of headings and
no code, tests or golden vectors, so nothing pins these restatements to the
reference's own numerics; they are the declared definition of correct
(BASELINE.md section 2).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import anything from this package.
"""
