"""ORACLE — TEST INFRASTRUCTURE ONLY.

Stand-in for the external ``utac_gym`` package (nanobind wrapper of the C++
``utac`` library) that the reference imports at
/root/reference/src/tacult/utac_game.py:2-3 and base/coach.py:17-18 but does
not vendor.  PARITY UNPINNED: see oracle/utac_rules.h.
"""
