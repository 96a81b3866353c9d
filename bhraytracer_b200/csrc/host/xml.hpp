// xml.hpp -- minimal XML DOM for the scene loader.
//
// The reference parses its scene files with a vendored tinyxml2 (reference
// BHRayTracer/Tinyxml2/, out of scope to re-ship).  The loader only needs: nested elements,
// quoted attributes, comments, the <?...?> / <!...> prologue, and tinyxml2's attribute query
// semantics (reference xmlload.cpp:478-521 via QueryDoubleAttribute / QueryIntAttribute:
// sscanf-style parse, caller's default kept when the attribute is missing or unparsable).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace bhrt { namespace host {

struct XmlElement {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlElement>> children;

    // tinyxml2 XMLElement::Attribute(name): exact (case-sensitive) match, nullptr if absent
    const char *Attribute(const char *n) const
    {
        for (auto &a : attrs) if (a.first == n) return a.second.c_str();
        return nullptr;
    }
    // QueryDoubleAttribute: true iff present and a number could be parsed (sscanf "%lf")
    bool QueryDouble(const char *n, double *out) const
    {
        const char *s = Attribute(n);
        if (!s) return false;
        char *end = nullptr;
        double v = strtod(s, &end);
        if (end == s) return false;
        *out = v;
        return true;
    }
    bool QueryInt(const char *n, int *out) const
    {
        const char *s = Attribute(n);
        if (!s) return false;
        char *end = nullptr;
        long v = strtol(s, &end, 10);
        if (end == s) return false;
        *out = (int)v;
        return true;
    }
    const XmlElement *FirstChildElement(const char *n) const
    {
        for (auto &c : children) if (c->name == n) return c.get();
        return nullptr;
    }
};

class XmlDocument {
public:
    std::unique_ptr<XmlElement> root;   // synthetic document node; top-level elements are its children
    std::string error;

    bool LoadFile(const char *path)
    {
        FILE *fp = fopen(path, "rb");
        if (!fp) { error = std::string("cannot open ") + path; return false; }
        std::string text;
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, n);
        fclose(fp);
        return Parse(text);
    }

    bool Parse(const std::string &text)
    {
        s_ = text.c_str();
        end_ = s_ + text.size();
        root.reset(new XmlElement);
        root->name = "#document";
        if (!ParseChildren(root.get(), nullptr)) return false;
        return true;
    }

private:
    const char *s_ = nullptr, *end_ = nullptr;

    static bool IsNameChar(char c) { return isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }
    void SkipWs() { while (s_ < end_ && isspace((unsigned char)*s_)) s_++; }
    bool StartsWith(const char *lit) const { size_t n = strlen(lit); return (size_t)(end_ - s_) >= n && strncmp(s_, lit, n) == 0; }
    bool SkipUntil(const char *lit)
    {
        size_t n = strlen(lit);
        while ((size_t)(end_ - s_) >= n) {
            if (strncmp(s_, lit, n) == 0) { s_ += n; return true; }
            s_++;
        }
        error = std::string("unterminated construct, expected ") + lit;
        return false;
    }
    static std::string Decode(const std::string &v)
    {
        std::string o;
        for (size_t i = 0; i < v.size(); i++) {
            if (v[i] == '&') {
                static const struct { const char *e; char c; } ents[] = { {"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''} };
                bool hit = false;
                for (auto &e : ents) {
                    size_t n = strlen(e.e);
                    if (v.compare(i, n, e.e) == 0) { o.push_back(e.c); i += n - 1; hit = true; break; }
                }
                if (hit) continue;
            }
            o.push_back(v[i]);
        }
        return o;
    }

    // parses siblings until the closing tag of `parentName` (or EOF for the document node)
    bool ParseChildren(XmlElement *parent, const char *parentName)
    {
        for (;;) {
            while (s_ < end_ && *s_ != '<') s_++;   // text content is ignored
            if (s_ >= end_) {
                if (parentName) { error = std::string("missing </") + parentName + ">"; return false; }
                return true;
            }
            if (StartsWith("<!--")) { if (!SkipUntil("-->")) return false; continue; }
            if (StartsWith("<?")) { if (!SkipUntil("?>")) return false; continue; }
            if (StartsWith("<![CDATA[")) { if (!SkipUntil("]]>")) return false; continue; }
            if (StartsWith("<!")) { if (!SkipUntil(">")) return false; continue; }
            if (StartsWith("</")) {
                s_ += 2;
                const char *b = s_;
                while (s_ < end_ && IsNameChar(*s_)) s_++;
                std::string nm(b, s_);
                if (!SkipUntil(">")) return false;
                if (!parentName || nm != parentName) { error = "mismatched closing tag </" + nm + ">"; return false; }
                return true;
            }
            // element
            s_++;
            const char *b = s_;
            while (s_ < end_ && IsNameChar(*s_)) s_++;
            if (s_ == b) { error = "bad element name"; return false; }
            std::unique_ptr<XmlElement> el(new XmlElement);
            el->name.assign(b, s_);
            bool selfClosed = false;
            for (;;) {
                SkipWs();
                if (s_ >= end_) { error = "unterminated tag <" + el->name; return false; }
                if (*s_ == '>') { s_++; break; }
                if (*s_ == '/') { s_++; SkipWs(); if (s_ < end_ && *s_ == '>') { s_++; selfClosed = true; break; } error = "bad '/' in tag"; return false; }
                const char *ab = s_;
                while (s_ < end_ && IsNameChar(*s_)) s_++;
                if (s_ == ab) { error = "bad attribute in <" + el->name + ">"; return false; }
                std::string an(ab, s_);
                SkipWs();
                if (s_ >= end_ || *s_ != '=') { error = "attribute without value: " + an; return false; }
                s_++;
                SkipWs();
                if (s_ >= end_ || (*s_ != '"' && *s_ != '\'')) { error = "unquoted attribute: " + an; return false; }
                char q = *s_++;
                const char *vb = s_;
                while (s_ < end_ && *s_ != q) s_++;
                if (s_ >= end_) { error = "unterminated attribute value: " + an; return false; }
                el->attrs.emplace_back(an, Decode(std::string(vb, s_)));
                s_++;
            }
            XmlElement *raw = el.get();
            parent->children.push_back(std::move(el));
            if (!selfClosed) {
                std::string nm = raw->name;
                if (!ParseChildren(raw, nm.c_str())) return false;
            }
        }
    }
};

}}  // namespace
