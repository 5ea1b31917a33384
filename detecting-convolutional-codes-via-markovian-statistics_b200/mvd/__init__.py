"""mvd -- host side of the B200 Markov-Viterbi detector (tables, bit sources, C-ABI binding)."""
