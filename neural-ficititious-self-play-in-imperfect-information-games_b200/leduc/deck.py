"""Host-side Card / Deck with the reference's interface (leduc/deck.py:5-55).

The batched engine never builds decks: it deals rank triples straight from Philox
(csrc/philox.cuh deal_ranks).  These classes exist for callers that construct decks themselves
and as the trace-injection format (a list of ranks in pop order)."""
import random

from .cardmatrix import RANKS, SUITS


class Card:
    def __init__(self, rank, suit):
        self._rank, self._suit = rank, suit
        self._named_rank, self._named_suit = RANKS[rank], SUITS[suit]

    def __str__(self):
        return "%s %s %s %s" % (self._named_rank, self._named_suit, self._rank, self._suit)

    @property
    def rank(self):
        return self._rank


class Deck:
    def __init__(self, size=6):
        if size <= 0 or size % 2:
            raise AssertionError("Decksize has to be an even number which is greater than 0.")
        self._size = size
        self._cards = [Card(r, s) for r in range(size // 2) for s in range(2)]
        self.fake_pub = Card(-1, -1)

    def shuffle(self):
        random.shuffle(self._cards)

    def fake_pub_card(self):
        return self.fake_pub

    def pick_up(self):
        return self._cards.pop()

    def print_deck(self):
        for c in self._cards:
            print(c)
