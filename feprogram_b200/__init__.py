"""feprogram_b200 -- B200-native FEM hot path (element matrices -> CSR assembly -> Jacobi-PCG)."""
__version__ = "0.1.0"
