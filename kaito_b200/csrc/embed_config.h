// embed_config.h -- shape of the BERT encoder served by K5 (shared by embed.cu and api.cu)
#pragma once
namespace krag {
struct BertConfig { int layers, hidden, heads, inter, vocab, max_pos, type_vocab; float eps; };
}
