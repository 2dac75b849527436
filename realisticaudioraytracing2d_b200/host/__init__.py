"""Host-side mirror of the reference's C# component surface for the hot path."""
