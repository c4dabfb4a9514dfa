"""excel_b200: B200-native (sm_100a) CAM -> SVC -> PAR hot path of ExCEL behind the reference's call surface."""
__version__ = "0.1.0"
