"""Host text pipeline: Snowball English stemmer against the algorithm's published sample
vocabulary, and the bm25s-style tokeniser (lower-case, \\w\\w+, 33 stop words)."""
from kaito_b200 import text

# word -> stem pairs from the Snowball English stemmer description (sample vocabulary)
SAMPLE = """consign consign consigned consign consigning consign consignment consign consist consist consisted consist
consistency consist consistent consist consistently consist consisting consist consists consist consolation consol
consolations consol consolatory consolatori console consol consoled consol consoles consol consolidate consolid
consolidated consolid consolidating consolid consoling consol consolingly consol consols consol consonant conson
consort consort consorted consort consorting consort conspicuous conspicu conspicuously conspicu conspiracy conspiraci
conspirator conspir conspirators conspir conspire conspir conspired conspir conspiring conspir constable constabl
constables constabl constance constanc constancy constanc constant constant knack knack knackeries knackeri knacks knack
knag knag knave knave knaves knave knavish knavish kneaded knead kneading knead knee knee kneel kneel kneeled kneel
kneeling kneel kneels kneel knees knee knell knell knelt knelt knew knew knick knick knif knif knife knife knight knight
knightly knight knights knight knit knit knits knit knitted knit knitting knit knives knive knob knob knobs knob knock knock
knocked knock knocker knocker knockers knocker knocking knock knocks knock knopp knopp knot knot knots knot""".split()

EXTRA = {"caresses": "caress", "ponies": "poni", "ties": "tie", "cries": "cri", "gas": "gas", "gaps": "gap", "feed": "feed",
         "agreed": "agre", "plastered": "plaster", "bled": "bled", "motoring": "motor", "sing": "sing", "hopping": "hop",
         "hoping": "hope", "happy": "happi", "sky": "sky", "dying": "die", "generously": "generous", "communism": "communism",
         "relational": "relat", "conditional": "condit", "rational": "ration", "generalization": "general",
         "running": "run", "documents": "document", "retrieval": "retriev", "databases": "databas", "queries": "queri",
         "proceed": "proceed", "succeed": "succeed", "y": "y", "by": "by", "say": "say", "is": "is"}


def test_snowball_sample_vocabulary():
    pairs = list(zip(SAMPLE[0::2], SAMPLE[1::2]))
    assert len(pairs) == 80
    bad = [(w, s, text.stem(w)) for w, s in pairs if text.stem(w) != s]
    assert not bad, bad


def test_snowball_rule_examples():
    bad = [(w, s, text.stem(w)) for w, s in EXTRA.items() if text.stem(w) != s]
    assert not bad, bad


def test_tokenize_bm25s_style():
    assert text.tokenize("The quick brown foxes are JUMPING over a lazy dog, and it's 2 o'clock!") == \
        ["quick", "brown", "fox", "jump", "over", "lazi", "dog", "clock"]
    assert text.tokenize("") == [] and text.tokenize("a I x") == []
    assert len(text.STOPWORDS_EN) == 33


def test_vocabulary_ids_and_query_terms():
    v = text.Vocabulary()
    ids, tf, dl = v.doc_terms("first document about documents. First!")
    assert dl == 5 and sorted(tf.tolist()) == [1, 2, 2]
    assert [v.terms[i] for i in ids] == ["first", "document", "about"]
    q = v.query_terms("what is the first FIRST document unknownword")
    assert [v.terms[i] for i in q] == ["first", "first", "document"]
