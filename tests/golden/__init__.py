"""Known-answer vectors transcribed from the reference's own gtests (file:line cited per case)."""
