"""Import the UNMODIFIED reference (Rintazero/ft8_demodulator) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py and by the tests marked
`needs_reference` (they skip when /root/reference is absent, e.g. on the GPU box).
The product package never imports this module.

Recipe follows SURVEY.md Appendix C: a no-op matplotlib shim (only if the real
matplotlib is missing), sys.path insert of /root/reference/src, stdout silenced,
and a writable scratch CWD for the PNG the reference writes (ft8_decode.py:379).
"""
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("FT8_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "ft8_tools"))


class Reference:
    """Handle to the reference's modules with quiet call helpers."""

    def __init__(self):
        if not available():
            raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
        if importlib.util.find_spec("matplotlib") is None:
            sys.path.insert(0, _SHIM)
        src = os.path.join(REFERENCE_ROOT, "src")
        if src not in sys.path:
            sys.path.insert(0, src)
        with contextlib.redirect_stdout(io.StringIO()):
            self.spectrogram_analyse = importlib.import_module(
                "ft8_tools.ft8_demodulator.spectrogram_analyse")
            self.ft8_decode = importlib.import_module("ft8_tools.ft8_demodulator.ft8_decode")
            self.ldpc_decoder = importlib.import_module("ft8_tools.ft8_demodulator.ldpc_decoder")
            self.crc = importlib.import_module("ft8_tools.ft8_demodulator.crc")
            self.ftx_types = importlib.import_module("ft8_tools.ft8_demodulator.ftx_types")
            self.constants = importlib.import_module("ft8_tools.ft8_demodulator.constants")
            self.generator = importlib.import_module("ft8_tools.ft8_generator")
            self.gen_encoder = importlib.import_module("ft8_tools.ft8_generator.encoder")
            self.gen_crc = importlib.import_module("ft8_tools.ft8_generator.crc")
            self.gen_ldpc = importlib.import_module("ft8_tools.ft8_generator.ldpc")
            self.gen_modulator = importlib.import_module("ft8_tools.ft8_generator.modulator")
            try:   # needs scikit-learn (present in the build image, absent from the reference's requirements.txt)
                self.frequency_correction = importlib.import_module("ft8_tools.ft8_beacon_receiver.frequency_correction")
                import logging
                logging.getLogger("ft8_tools.ft8_beacon_receiver.frequency_correction").setLevel(logging.CRITICAL)
                self.frequency_correction.logger.setLevel(logging.CRITICAL)
            except Exception:   # pragma: no cover
                self.frequency_correction = None
        self._scratch = tempfile.mkdtemp(prefix="ft8ref_")

    def quiet(self, fn, *args, **kwargs):
        """Call a reference function with stdout silenced, from a scratch CWD."""
        cwd = os.getcwd()
        os.chdir(self._scratch)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                return fn(*args, **kwargs)
        finally:
            os.chdir(cwd)


_ref = None


def load() -> Reference:
    global _ref
    if _ref is None:
        _ref = Reference()
    return _ref
