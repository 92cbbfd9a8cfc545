"""b200pinn -- B200-native PINN hot path."""
