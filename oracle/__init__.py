"""CPU oracle for the EnSRF hot path -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  See oracle/ensrf_oracle.py for the restatement and its pinning.  The reference is pure Python
(no C/C++ sources), so there is no oracle/_ref build.
"""
