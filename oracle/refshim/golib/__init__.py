"""Minimal stand-in for the author's un-vendored Golib package (TEST INFRASTRUCTURE ONLY).

The reference (ArnaudPel/CamKifu) imports `golib` for the goban constants and the Move type; Golib is a sister
project that is not part of /root/reference. This shim provides just the names the stone-detection hot path touches
so that the *unmodified* reference modules can be imported in the authoring container to generate golden vectors
(see oracle/gen_golden.py). Nothing in the product path imports this.
"""
