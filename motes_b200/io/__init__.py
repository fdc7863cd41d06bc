"""Host-side I/O either side of the extraction path: a FITS shim with the slice of the
astropy API MOTES uses (``fits``), the harvest / save functions of ``motes.io.motesio`` and ``motes.io.gmosio``
(``motesio``, ``gmosio``) and ``standin.install()``, which lets the unmodified reference import where astropy
is missing.  Nothing here computes on the extraction path."""
