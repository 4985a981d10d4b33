"""Stub so that camkifu.stone.nn_manager (which imports keras at module top) can be imported for its geometry and
label codec. The network itself is never built through this stub. TEST INFRASTRUCTURE ONLY."""
