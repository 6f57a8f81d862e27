"""CPU oracle (test infrastructure).  See unet_oracle.py / schedule_oracle.py headers."""
