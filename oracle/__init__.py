"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU (torch fp32 / numpy float64) restatement of the reference's algorithm for the denoising hot path
(SURVEY.md section 8a), used as the parity checker.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it; the product package
(kandinsky-2_b200/kandinsky2) never does -- it fails loudly if libk2b200.so is missing.

Pinning: the reference ships no tests or golden vectors, so the restatement is pinned
against the reference's own code executed in the build container (oracle/ref_shim.py imports
/root/reference/kandinsky2 through a namespace stub; oracle/make_golden.py writes tests/golden/*.pt).
The 2.2 conditioning head (diffusers, not in /root/reference) is "parity unpinned" -- see DESIGN.md.
"""
