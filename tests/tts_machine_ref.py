"""Test-side restatement of the reference's TTS text scheduling (src/moshi/models/lm.h:5-194: TokenIds, State,
StateMachine::process), written independently of the C++ in moshi.cpp_b200/host/moshi_api.cpp so that the two can be
checked against each other."""
from collections import deque

NEW_WORD, PAD, ZERO = 0, 3, -1          # lm.h:8-12


class Machine:
    def __init__(self, card, second_stream_ahead=0, max_padding=6, initial_padding=2):
        self.card, self.ahead, self.max_padding, self.initial_padding = card, second_stream_ahead, max_padding, initial_padding
        self.reset()

    def reset(self):                    # lm.h:95-102
        self.remaining_padding = self.initial_padding
        self.forced_padding = self.initial_padding
        self.end_step = -1
        self.entries, self.queued, self.lookahead = deque(), deque(), deque()

    def push(self, tokens, padding):
        self.entries.append((list(tokens), padding))

    def is_empty(self):                 # lm.h:41-49
        return not (self.entries or self.queued or self.lookahead)

    def _tokens_ahead(self, lookahead):  # lm.h:28-39
        for tokens, _ in self.entries:
            if not tokens:
                continue
            lookahead -= 1
            if lookahead != 0:
                continue
            return tokens
        return []

    def process(self, step, token):     # lm.h:104-193
        if token not in (NEW_WORD, PAD):
            token = PAD
        if self.queued:
            token = PAD
        elif self.forced_padding > 0:
            token = PAD
        elif self.remaining_padding <= 0:
            token = NEW_WORD
        if token == NEW_WORD:
            if self.entries:
                tokens, padding = self.entries.popleft()
                if tokens:
                    self.queued.extend(tokens)
                    if self.ahead:
                        self.lookahead.extend(self._tokens_ahead(self.ahead))
                    self.remaining_padding = self.max_padding
                else:
                    token = PAD
                self.forced_padding = padding
            else:
                token = PAD
                if self.ahead and self.end_step < 0:
                    token = NEW_WORD
                if self.end_step < 0:
                    self.end_step = step
        output = 0
        if token == PAD:
            if self.remaining_padding > 0:
                self.remaining_padding -= 1
            if self.forced_padding > 0:
                self.forced_padding -= 1
            output = self.queued.popleft() if self.queued else PAD
        elif token == NEW_WORD:
            output = NEW_WORD
        elif token == ZERO:
            output = token
        if self.ahead:
            second = -1
            if output == NEW_WORD:
                second = NEW_WORD
                output = self.queued.popleft() if self.queued else PAD
            elif self.lookahead:
                second = self.lookahead.popleft()
            output = (second + 1) * self.card + output
        return output
