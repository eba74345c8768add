"""TEST INFRASTRUCTURE ONLY.  CPU/fp32 restatement of the reference MMDiT path, used as the
parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
Nothing under stable-diffusion-3-from-scratch_b200/ imports this package."""
