"""TEST INFRASTRUCTURE — parity checkers for the PanDelos Pangenes hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker or the timed CPU baseline.

* ``oracle.cport``      ctypes binding of ``pangenes_oracle.c`` (plain-C restatement, builds anywhere)
* ``oracle.refjni``     ctypes binding of ``fakejni.cpp``: drives any ``libnative.so``-shaped library (the unmodified
                        reference build ``oracle/_ref/libnative_ref.so`` or the B200 drop-in) through a fake JNIEnv
* ``oracle.pangenes_java``  numpy restatement of the Java side of the path (``PangeneIData``, the BBH filter of
                        ``Pangenes.main``, ``PangeneNet.saveToFile``) and of ``calculate_k.py``
"""
