"""ft8_demodulator_b200 - B200-native FT8 demodulation hot path."""
__version__ = "0.1.0"
