"""TEST INFRASTRUCTURE ONLY.

CPU restatements (torch-CPU fp32 / numpy f64) of the pero-ocr line-recognition
hot path, used as the parity checker.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package.  Nothing under ``pero_ocr_b200/`` imports it.
"""
