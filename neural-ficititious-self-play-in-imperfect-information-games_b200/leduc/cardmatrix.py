"""Card display names (reference leduc/cardmatrix.py:4-12).  The reference builds a numpy matrix
per Card (7 per reset, pure overhead); here it is a constant table."""

RANKS = ("Ace", "King", "Queen", "Jack", "10", "9", "8", "7", "6", "5", "4", "3", "2")
SUITS = ("Heart", "Spades", "Cross", "Diamonds")


class Cardmatrix:
    def getCard(self, rank, suit):
        return RANKS[rank], SUITS[suit]
