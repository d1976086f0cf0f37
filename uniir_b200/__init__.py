"""uniir_b200: B200-native (sm_100a) hot path of TIGER-AI-Lab/UniIR behind the reference's
own model / embed / index / retrieval interfaces.  See DESIGN.md."""
__version__ = "0.1.0"
