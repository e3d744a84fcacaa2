// foamDictionary.hpp — a minimal reader of OpenFOAM dictionary files, enough for the case files the reference's
// beamFoam reads for this path (constant/beamProperties, constant/g, 0/W, 0/Theta, system/controlDict,
// system/fvSchemes and the (time value) table files of the boundary-condition series).  Syntax handled: C and C++
// comments, `key value...;`, `key { ... }`, quoted and "a|b" alternation keys, lists `( ... )` incl. lists of named
// dictionaries (`beams ( beam_0 { ... } )`), dimension sets `[...]` (ignored), $FOAM_CASE substitution.
#pragma once
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace foam
{
struct Dict;
struct Value {
    enum Kind { NUMBER, WORD, STRING, LIST, NAMED_DICT } kind = WORD;
    double v = 0.0;
    std::string s;                 // WORD / STRING text, or the name of a NAMED_DICT
    std::vector<Value> list;       // LIST elements
    std::shared_ptr<Dict> dict;    // NAMED_DICT body
};
struct Entry {
    std::string key;
    std::vector<Value> values;     // tokens up to ';'
    std::shared_ptr<Dict> dict;    // non-null: `key { ... }`
};
struct Dict {
    std::vector<Entry> entries;
    std::string file;

    const Entry* find(const std::string& name) const
    {
        for (const Entry& e : entries) { // the LAST match wins in OpenFOAM; files here have no duplicates
            size_t p = 0;
            while (p <= e.key.size()) {   // "fileName|file" style alternation
                size_t q = e.key.find('|', p);
                if (q == std::string::npos) q = e.key.size();
                if (e.key.compare(p, q - p, name) == 0) return &e;
                p = q + 1;
            }
        }
        return nullptr;
    }
    bool found(const std::string& k) const { return find(k) != nullptr; }
    const Entry& lookup(const std::string& k) const
    {
        const Entry* e = find(k);
        if (!e) throw std::runtime_error("keyword " + k + " is undefined in dictionary " + file);
        return *e;
    }
    const Dict& subDict(const std::string& k) const
    {
        const Entry& e = lookup(k);
        if (!e.dict) throw std::runtime_error("keyword " + k + " is not a sub-dictionary in " + file);
        return *e.dict;
    }
    // scalar or dimensionedScalar (`E E [1 -1 -2 0 0 0 0] 7.9e4;`): the last number of the entry
    double scalar(const std::string& k) const
    {
        const Entry& e = lookup(k);
        for (size_t i = e.values.size(); i-- > 0;)
            if (e.values[i].kind == Value::NUMBER) return e.values[i].v;
        throw std::runtime_error("keyword " + k + " has no scalar value in " + file);
    }
    double scalarOr(const std::string& k, double d) const { return found(k) ? scalar(k) : d; }
    std::string word(const std::string& k) const
    {
        const Entry& e = lookup(k);
        if (e.values.empty()) throw std::runtime_error("keyword " + k + " has no value in " + file);
        return e.values[0].s;
    }
    std::string wordOr(const std::string& k, const std::string& d) const { return found(k) ? word(k) : d; }
    // `key (x y z);` or `key uniform (x y z);`
    std::vector<double> vector(const std::string& k) const
    {
        const Entry& e = lookup(k);
        for (const Value& v : e.values)
            if (v.kind == Value::LIST) {
                std::vector<double> r;
                for (const Value& c : v.list) r.push_back(c.v);
                return r;
            }
        throw std::runtime_error("keyword " + k + " has no vector value in " + file);
    }
};

namespace detail
{
struct Tok { char p = 0; bool quoted = false; std::string s; }; // p != 0: punctuation
inline std::vector<Tok> tokenize(const std::string& t)
{
    std::vector<Tok> out;
    size_t i = 0, n = t.size();
    while (i < n) {
        const char c = t[i];
        if (isspace((unsigned char)c)) { i++; continue; }
        if (c == '/' && i + 1 < n && t[i + 1] == '/') { while (i < n && t[i] != '\n') i++; continue; }
        if (c == '/' && i + 1 < n && t[i + 1] == '*') { i += 2; while (i + 1 < n && !(t[i] == '*' && t[i + 1] == '/')) i++; i += 2; continue; }
        if (c == '{' || c == '}' || c == '(' || c == ')' || c == ';' || c == '[' || c == ']') { Tok k; k.p = c; out.push_back(k); i++; continue; }
        Tok k;
        if (c == '"') { i++; while (i < n && t[i] != '"') k.s += t[i++]; i++; k.quoted = true; }
        else while (i < n && !isspace((unsigned char)t[i]) && !strchr("{}();[]\"", t[i])) k.s += t[i++];
        out.push_back(k);
    }
    return out;
}
inline Value atom(const Tok& k)
{
    Value v;
    if (k.quoted) { v.kind = Value::STRING; v.s = k.s; return v; }
    char* end = nullptr;
    const double d = strtod(k.s.c_str(), &end);
    if (end && *end == 0 && !k.s.empty()) { v.kind = Value::NUMBER; v.v = d; v.s = k.s; }
    else { v.kind = Value::WORD; v.s = k.s; }
    return v;
}
inline std::shared_ptr<Dict> parseDict(const std::vector<Tok>& t, size_t& i, bool nested, const std::string& file);
inline Value parseList(const std::vector<Tok>& t, size_t& i, const std::string& file) // t[i-1] was '('
{
    Value L;
    L.kind = Value::LIST;
    while (i < t.size() && t[i].p != ')') {
        if (t[i].p == '(') { i++; L.list.push_back(parseList(t, i, file)); continue; }
        if (t[i].p == ';') { i++; continue; }
        if (t[i].p) throw std::runtime_error("unexpected '" + std::string(1, t[i].p) + "' in a list in " + file);
        if (i + 1 < t.size() && t[i + 1].p == '{') { // named dictionary element
            Value v;
            v.kind = Value::NAMED_DICT; v.s = t[i].s;
            i += 2;
            v.dict = parseDict(t, i, true, file);
            L.list.push_back(v);
            continue;
        }
        L.list.push_back(atom(t[i++]));
    }
    if (i >= t.size()) throw std::runtime_error("unterminated list in " + file);
    i++; // ')'
    return L;
}
inline std::shared_ptr<Dict> parseDict(const std::vector<Tok>& t, size_t& i, bool nested, const std::string& file)
{
    auto d = std::make_shared<Dict>();
    d->file = file;
    while (i < t.size()) {
        if (t[i].p == '}') { if (!nested) throw std::runtime_error("unbalanced '}' in " + file); i++; return d; }
        if (t[i].p == ';') { i++; continue; }
        if (t[i].p) throw std::runtime_error("unexpected '" + std::string(1, t[i].p) + "' where a keyword should be in " + file);
        Entry e;
        e.key = t[i++].s;
        if (i < t.size() && t[i].p == '{') { i++; e.dict = parseDict(t, i, true, file); d->entries.push_back(e); continue; }
        while (i < t.size() && t[i].p != ';') {
            if (t[i].p == '(') { i++; e.values.push_back(parseList(t, i, file)); }
            else if (t[i].p == '[') { while (i < t.size() && t[i].p != ']') i++; i++; } // dimension set
            else if (t[i].p == '}' ) break; // tolerate a missing ';' before '}'
            else if (t[i].p) throw std::runtime_error("unexpected '" + std::string(1, t[i].p) + "' in the value of " + e.key + " in " + file);
            else e.values.push_back(atom(t[i++]));
        }
        if (i < t.size() && t[i].p == ';') i++;
        d->entries.push_back(e);
    }
    if (nested) throw std::runtime_error("unterminated '{' in " + file);
    return d;
}
inline std::string slurp(const std::string& path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
} // namespace detail

inline std::shared_ptr<Dict> readDictionary(const std::string& path)
{
    const auto toks = detail::tokenize(detail::slurp(path));
    size_t i = 0;
    return detail::parseDict(toks, i, false, path);
}
// a file that holds one bare list, e.g. the (time value) tables of the BC series
inline Value readList(const std::string& path)
{
    const auto toks = detail::tokenize(detail::slurp(path));
    size_t i = 0;
    while (i < toks.size() && toks[i].p != '(') i++;
    if (i >= toks.size()) throw std::runtime_error("no list found in " + path);
    i++;
    return detail::parseList(toks, i, path);
}
inline std::string expand(std::string s, const std::string& caseDir)
{
    for (const char* var : {"$FOAM_CASE", "${FOAM_CASE}"}) {
        size_t p;
        while ((p = s.find(var)) != std::string::npos) s.replace(p, strlen(var), caseDir);
    }
    return s;
}
} // namespace foam
