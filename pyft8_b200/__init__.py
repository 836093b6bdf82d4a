"""pyft8_b200 -- B200-native (sm_100a CUDA) implementation of G1OJS/PyFT8's receive hot path.

audio -> waterfall -> Costas sync -> candidates -> LLRs -> fine sync -> LDPC(174,91) -> OSD -> CRC-14,
behind the reference's receiver.py / decoders.py call surface.  See DESIGN.md and include/ft8_b200.h.
"""
__all__ = ["engine", "decoders", "receiver", "messages", "synth", "tables"]
