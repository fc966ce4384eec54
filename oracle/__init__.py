"""CPU oracle for the DSP-AudioRecLabs front end + KNN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dsp_audioreclabs_b200/`` may import,
link or execute anything in this package: the product path is the CUDA
library (``libdspfront.so``) and it fails loudly when that library is absent.
The only permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: PINNED.  The restatement in ``frontend_oracle.py`` /
``knn_oracle.py`` / ``frontend_oracle.c`` is checked (tests/test_oracle_golden.py)
against fixtures in ``tests/golden/`` that were produced by importing the
reference's own ``src.audio_processing`` / ``src.feature_extraction`` /
``src.models`` from a copy of /root/reference (generator: oracle/gen_golden.py).
The reference itself ships no tests or golden vectors.
"""
