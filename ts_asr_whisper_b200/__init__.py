"""Importable alias of the hyphenated source directory ``ts-asr-whisper_b200/``.

Python identifiers cannot contain '-', so this shim package points its ``__path__`` at the real source directory;
``import ts_asr_whisper_b200.ops`` then loads ``ts-asr-whisper_b200/ops.py``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ts-asr-whisper_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
