"""CPU oracle for the RecBox hot path -- TEST INFRASTRUCTURE ONLY (see recbox_oracle.py header).
Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by recbox_b200/."""
