"""CPU oracle for the full-sky Gaussian field path of radiocosmology/cora.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cora_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU arm.

Every function is a restatement (numpy / scipy / a few lines of C) of the reference
algorithm, citing the reference ``file:line`` it follows.  Parity status:

* C_l stage (``spectra.py``, ``skysim.py:clarray``): PINNED against the reference's
  own known-answer values (``tests/test_corr.py:18,28-29,44,54-55``) and against
  fixtures produced by importing the real reference (``tests/golden/make_golden.py``).
* root / draws / apply (``nputil.py``, ``skysim.py:mkfullsky``): PINNED against
  fixtures produced by running the real reference ``mkfullsky(alms=True)``.
* inverse SHT (``sht.py``): the arithmetic lives in healpy/libsharp (third party,
  ``healpy>=1.17`` per ``pyproject.toml:30``), which is neither under /root/reference
  nor installable here, and the reference's tests hold no golden at that boundary:
  **parity unpinned**.  ``sht.py`` restates the published HEALPix RING synthesis and is
  validated analytically (scipy ``sph_harm_y`` direct sums, closed forms, and exact rational
  evaluations of the spin-0 / spin-2 closed sums for every (l, m) up to l = 64).
* forward SHT (``sht.py: map2alm / anafast``; ``healpy.map2alm``): **parity unpinned** for the
  same reason; pinned to the synthesis through the adjoint identity and a band-limited round trip.
"""
