"""Mirror of /root/reference/callbacks (EMA weight averaging, main.py:131 ``--ema``)."""
