"""CPU oracle for the DD-QST generative-tomography hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may.  See ``oracle/ddqst_oracle.py`` for the header
that states how the restatement is pinned.
"""
